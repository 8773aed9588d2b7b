#!/bin/bash
# round-2 rehearsal of what the driver runs at round end (1 GPU): full GPU suite, smoke(), both bench arms; then the profile
# captures profiles/README.md cites (ncu --set full of the selected K5 sweep, launch list of the bench command, dram bytes of
# the other K5 kernels).   gpurun --timeout 2400 -- bash scripts/gpu_r2_final.sh
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2z_pytest.log 2>&1; grep -E "passed|failed|rror" gpurun_out/r2z_pytest.log | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --impl reference ) > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; cut -c1-200 gpurun_out/r2z_bench_ref.json
( time python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; cut -c1-200 gpurun_out/r2z_bench.json; tail -2 gpurun_out/r2z_bench.err
# profiles
timeout 600 ncu --set full --clock-control none --import-source on --kernel-id ::ps_sweep:70 -f -o gpurun_out/r02_ps_sweep python scripts/prof_plane.py 20 > gpurun_out/r02_prof.log 2>&1; tail -1 gpurun_out/r02_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-stable --no-extras --no-cpu-baseline --no-converge > gpurun_out/r02_bench_under_ncu.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pi_build_rows|improve_kernel|compact_rows_kernel|plane_items_kernel|plane_slots_kernel|gp_sweep" -s 160 -c 14 --csv --log-file gpurun_out/r02_k5_kernels.csv python scripts/prof_plane.py 20 > /dev/null 2>&1
tail -16 gpurun_out/r02_k5_kernels.csv | cut -c1-200
