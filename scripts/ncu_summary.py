#!/usr/bin/env python3
"""Text summary of one `ncu --set full --import-source on` capture (a .ncu-rep with one kernel launch), CPU only:
selected raw metrics, stall reasons, the most sampled SASS instructions, shared-memory wavefronts per load opcode.
    python scripts/ncu_summary.py gpurun_out/r02_ps_sweep.ncu-rep > profiles/...txt"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for k in want:
    if k in M: print("   %-84s %s %s" % (k, M[k][0], M[k][1]))
print("   stall reasons (warps per issue-active cycle):")
for h in sorted(M):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        try:
            v = float(M[h][0])
        except ValueError:
            continue
        if v >= 0.1: print("      %-60s %.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
src = page("source")
hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(src[hi])}
data = src[hi + 1:]
def num(r, k):
    try:
        return int(r[ix[k]] or 0)
    except (ValueError, KeyError):
        return 0
tot = sum(num(r, "# Samples") for r in data)
print("   warp-state samples: %d over %d SASS instructions; the ten most sampled:" % (tot, len(data)))
for n, r in sorted(sorted(enumerate(data), key=lambda t: -num(t[1], "# Samples"))[:10]):
    print("      #%-5d %-64s %6d samples (%.1f %%), executed %d times" % (n, r[ix["Source"]][:64], num(r, "# Samples"), 100.0 * num(r, "# Samples") / max(tot, 1), num(r, "Instructions Executed")))
w = collections.Counter(); wi = collections.Counter(); ne = collections.Counter()
for r in data:
    toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
    if not toks: continue
    w[toks[0]] += num(r, "L1 Wavefronts Shared"); wi[toks[0]] += num(r, "L1 Wavefronts Shared Ideal"); ne[toks[0]] += num(r, "Instructions Executed")
print("   shared-memory wavefronts per opcode (measured / ideal / warp instructions):")
for k, v in w.most_common():
    if v: print("      %-10s %12d / %12d / %12d" % (k, v, wi[k], ne[k]))
print("   polling: SYNCS.PHASECHK.TRANS64.TRYWAIT executed %d times, BRA %d times" % (ne["SYNCS.PHASECHK.TRANS64.TRYWAIT"], ne["BRA"]))
