#!/bin/bash
# item-mode plane sweep: bit-parity tests on small grids, then K5 timing vs the other modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plane.py -x -q -m gpu -k "bit_identical or falls_back" > gpurun_out/r2b_plane_tests.log 2>&1
tail -3 gpurun_out/r2b_plane_tests.log
timeout 400 python scripts/exp_plane.py 20 "$@" > gpurun_out/r2b_exp_plane.log 2>&1
grep -v "^layout" gpurun_out/r2b_exp_plane.log | grep -v INFO | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d.get('cfg'), 'ms %.4f base %.4f mism %s loads %.1f late %.2f pairs %.1f stray %.2f singles-paired %.1f regs %s slots %s' % (d.get('ms_plane', 0), d.get('ms_base', 0), d.get('mismatches'), d.get('loads_per_plane', 0), d.get('late_per_plane', 0), d.get('pairs_per_plane', 0), d.get('stray_pairs_per_plane', 0), d.get('singles_paired_per_plane', 0), d.get('registers'), d.get('slots')), d.get('error', ''))
    else: print(l[:200])
"
