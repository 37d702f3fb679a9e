#!/bin/bash
# The round's measurements in one gpurun call:  gpurun --timeout 1500 -- 'bash scripts/profile_round.sh [tag]'
# Bench lines first (never under a profiler), then the ncu launch list of the same command and one
# `--set full` capture per hot kernel, summarised into gpurun_out/ (copy what should be judged to profiles/).
R=gpurun_out
T=${1:-r2}
NCU="ncu --set full --clock-control none --import-source on"
python bench.py > $R/${T}_bench_default.json 2> $R/${T}_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > $R/${T}_bench_reference.json 2>> $R/${T}_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $R/${T}_launches_contacts.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $R/ncu_launches.log 2>&1
python scripts/launch_summary.py $R/${T}_launches_contacts.csv "python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline" > $R/${T}_launches_contacts_summary.txt
$NCU -k regex:cell_sweep_kernel -s 2 -c 1 -f -o /tmp/sweep python scripts/bench_sweep.py 296 1 > $R/ncu_sweep.log 2>&1
python scripts/ncu_summary.py /tmp/sweep.ncu-rep cell_sweep > $R/${T}_ncu_cell_sweep_sym.txt
$NCU -k regex:stream_fused_kernel -s 2 -c 1 -f -o /tmp/rmsd python scripts/profile_fused.py rmsd 2048 > $R/ncu_rmsd.log 2>&1
python scripts/ncu_summary.py /tmp/rmsd.ncu-rep stream_fused > $R/${T}_ncu_rmsd_fused.txt
$NCU -k regex:stream_fused_kernel -s 2 -c 1 -f -o /tmp/rg python scripts/profile_fused.py rg 2048 > $R/ncu_rg.log 2>&1
python scripts/ncu_summary.py /tmp/rg.ncu-rep stream_fused > $R/${T}_ncu_rg_fused.txt
$NCU -k 'regex:rmsd_multi_(cov_ring|gradient_pk)' -s 4 -c 2 -f -o /tmp/multi python scripts/bench_path.py 1024 1 > $R/ncu_multi.log 2>&1
python scripts/ncu_summary.py /tmp/multi.ncu-rep rmsd_multi > $R/${T}_ncu_rmsd_multi_mma.txt
tail -c 600 $R/${T}_bench_default.json | head -c 300; echo; head -c 400 $R/${T}_bench_reference.json; echo
grep -c cell_sweep $R/${T}_launches_contacts.csv; head -4 $R/${T}_ncu_cell_sweep_sym.txt; head -4 $R/${T}_ncu_rmsd_fused.txt; head -4 $R/${T}_ncu_rg_fused.txt; head -3 $R/${T}_ncu_rmsd_multi_mma.txt
python scripts/bench_trajectory.py --frames 512 --chunk 64 > $R/${T}_bench_trajectory.json 2>> $R/${T}_bench_default.err
python scripts/bench_trajectory.py --frames 512 --chunk 64 --format xtc >> $R/${T}_bench_trajectory.json 2>> $R/${T}_bench_default.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $R/${T}_smoke.log 2>&1; tail -1 $R/${T}_smoke.log
