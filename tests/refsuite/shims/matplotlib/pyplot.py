class _Anything(object):
    def __getattr__(self, _name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


cm = _Anything()


def __getattr__(_name):  # figure, subplot, imshow, colorbar, xlabel, ...
    return _Anything()
