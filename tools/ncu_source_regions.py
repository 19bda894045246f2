#!/usr/bin/env python
"""Where a kernel's instructions and stall samples go, by source region and by 1 KB block of SASS.

Reads the source page of an .ncu-rep captured with `--set full --import-source on` (on the CPU box):

    python tools/ncu_source_regions.py gpurun_out/prof.ncu-rep 'title' [file.cu] > profiles/x.regions.txt

A region starts at the first source line matching one of MARKERS (the text comes from the report itself, so line
numbers never have to be kept in sync with the tree; only lines that carry SASS are listed there, so markers are
executable lines) and runs to the next marker.  An instruction inlined through several source lines is counted on
each of them, so the shares are relative to the per-line totals, not to the kernel's instruction count.  Lines of
other files (inlined CUDA headers) are listed per file.  The second table maps the SASS address space in 1 KB
blocks: share of executed instructions, of all stall samples and of the `no_instruction` samples -- the footprint
the instruction cache sees.
"""
import collections
import csv
import re
import subprocess
import sys

# (label, regex on the source text) in file order of treegp_b200/csrc/pairbin.cu
MARKERS = [
    ("bin search helpers (classification)", r"int i = \(int\)\(\(d \+ hi\) \* inv_bin\);|const int mid = \(lo \+ hi \+ 1\) >> 1;"),
    ("warp reductions", r"v = fmin\(v, __shfl_xor_sync"),
    ("per-lane coordinate thresholds (window open)", r"const long long b = __double_as_longlong\(d\);|return b >= 0 \? b :"),
    ("global rank query (two-axis form)", r"const int b = 8 \* \(\(p7 < T"),
    ("classify_twod", r"const double dx0 = cminx - imaxx, dx1 = cmaxx - iminx;"),
    ("kernel prologue", r"const int nb = P.nb, nbins = P.nbins, nwarps = P.warps;|const int tid = threadIdx.x"),
    ("flush_regs", r"const unsigned n_in = warp_sum_u\(A.nin\);"),
    ("fix_mirror", r"if \(!__any_sync\(0xffffffffu, A.mmc != 0u\)\) return;"),
    ("flush_hist", r"const unsigned c = my_c\[b\];|for \(int b = lane; b < nb; b \+= 32\) \{"),
    ("generic_block", r"for \(int jj = max\(j0, jfirst\)"),
    ("rank_query (shared memory)", r"const int b = 8 \* \(\(cxy\[7\]\.x < T"),
    ("cf_spill (closed-form bookings)", r"const int om = nb - 1 - cf_bin;|atomicAdd\(my_c \+ cf_bin, cf_cnt\);"),
    ("item decode + row load", r"if \(lane == 0\) q = atomicAdd\(P.counter, 1ull\);"),
    ("classification loop", r"const int64_t mychunk = sc \+ lane;|for \(int64_t sc = c_lo"),
    ("fetch / stage / prefetch", r"const unsigned todo = "),
    ("short cut: two-axis (quadrant) form", r"if \(\(\(fw >> 1\) & 3\) == 3\)"),
    ("short cut: one-axis rank query", r"const int dvar = "),
    ("general dispatch", r"const int64_t j0g = \(sc \+ c\) \* PB_CHUNK;"),
    ("window open", r"if \(!fits && !gen\) \{"),
    ("generic fallback call", r"^\s*if \(gen\) \{"),
    ("REG_FULL general (closed / 1-D / rank)", r"if \(bcls == PB_REG_FULL\) \{"),
    ("REG_FULL pair loops", r"PB_PAIR_NM_NT\(A, pj.x, pj.y, kj\);|PB_PAIR_NT\(A, pj.x, pj.y, kj\);"),
    ("REG_CHECK pair loop", r"const bool ok = r2 >= lo2 && fabs\(dx\) < M"),
    ("end of item / kernel tail", r"^\s*flush_hist\(cur_cat\);"),
]


def integer(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep, title = sys.argv[1], sys.argv[2]
    main_file = sys.argv[3] if len(sys.argv) > 3 else "pairbin.cu"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    i_ex, i_smp, i_noi = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("stall_no_inst")
    src, sass, cur_file, cur_line = [], {}, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0].isdigit():
            cur_line = (cur_file, int(r[0]), r[1])
            src.append(cur_line + (integer(r[i_ex]), integer(r[i_smp]), integer(r[i_noi])))
        elif r[0] == "" and len(r) > 3 and r[2].startswith("0x"):
            sass.setdefault(int(r[2], 16), (integer(r[i_ex]), integer(r[i_smp]), integer(r[i_noi])))
    t_ex, t_smp, t_noi = (max(1, sum(s[k] for s in src)) for k in (3, 4, 5))
    print(title)
    print("report: %s; totals: %.3g warp instructions, %d stall samples, %.1f %% of them no_instruction\n"
          % (rep, t_ex, t_smp, 100.0 * t_noi / t_smp))

    # ---- regions of the main file ----
    starts, found = [], set()
    for f, line, text, *_ in sorted((s for s in src if s[0] == main_file), key=lambda s: s[1]):
        for label, rx in MARKERS:
            if label not in found and re.search(rx, text):
                starts.append((line, label))
                found.add(label)
    starts.sort()
    acc = collections.OrderedDict()
    for f, line, text, ex, smp, noi in src:
        if f == main_file:
            label = "(before the first marker)"
            for ln, lb in starts:
                if line >= ln:
                    label = lb
        else:
            label = "[inlined] " + str(f)
        a = acc.setdefault(label, [0, 0, 0])
        a[0] += ex
        a[1] += smp
        a[2] += noi
    print("%-52s %8s %8s %8s" % ("region of " + main_file, "inst %", "samples %", "no_inst %"))
    order = [lb for _, lb in starts]
    for label in sorted(acc, key=lambda lb: (order.index(lb) if lb in order else 10 ** 6, lb)):
        ex, smp, noi = acc[label]
        print("%-52s %8.1f %8.1f %8.1f" % (label[:52], 100.0 * ex / t_ex, 100.0 * smp / t_smp, 100.0 * noi / t_noi))

    # ---- SASS footprint ----
    addrs = sorted(sass)
    base = addrs[0]
    blocks = collections.OrderedDict()
    for a in addrs:
        b = blocks.setdefault((a - base) // 1024, [0, 0, 0])
        for k in range(3):
            b[k] += sass[a][k]
    s_ex, s_smp, s_noi = (max(1, sum(v[k] for v in sass.values())) for k in range(3))
    warm = sum(1 for b in blocks.values() if b[0] > 0)
    hot = sum(1 for b in blocks.values() if b[0] > 0.004 * s_ex)
    print("\nSASS: %d instructions = %.0f KB; %d KB executed at all, %d KB with > 0.4 %% of the executed instructions"
          % (len(addrs), len(addrs) * 16 / 1024.0, warm, hot))
    print("%-6s %8s %8s %8s" % ("KB", "inst %", "samples %", "no_inst %"))
    for kb, (ex, smp, noi) in blocks.items():
        if ex:
            print("%-6d %8.2f %8.2f %8.2f" % (kb, 100.0 * ex / s_ex, 100.0 * smp / s_smp, 100.0 * noi / s_noi))


if __name__ == "__main__":
    main()
