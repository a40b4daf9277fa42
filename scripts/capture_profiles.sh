#!/bin/bash
# Run on the GPU box (gpurun -- 'bash scripts/capture_profiles.sh r02x'): the ncu captures and bench lines that
# profiles/ is built from.  Read back here with scripts/ncu_summarize.py.  PBA_NO_GRAPH=1 where K_B / in-loop K_A must be
# visible: ncu cannot profile kernel nodes inside a graph that holds a conditional (WHILE) node.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
DM=sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,sm__ops_path_tensor_src_fp64.sum,smsp__inst_executed_pipe_fp64.sum
python -c "
import numpy as np, sys
sys.path.insert(0, '.')
from workloads import synthetic
np.save('/tmp/cfg3_images.npy', synthetic.make_window().images)"
PBA_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cfg4 > $out/${tag}_launches_bench.log 2>&1
timeout 600 ncu --set full --metrics $DM --clock-control none --import-source on -k regex:k_step -s 2 -c 1 -f -o $out/${tag}_kstep \
    python scripts/prof_k1.py > $out/${tag}_kstep.log 2>&1
PBA_NO_GRAPH=1 timeout 600 ncu --set full --metrics $DM --clock-control none --import-source on -k regex:k_schur_solve -s 3 -c 1 -f -o $out/${tag}_kschur \
    python scripts/prof_k1.py solve > $out/${tag}_kschur.log 2>&1
PBA_NO_GRAPH=1 timeout 600 ncu --set full --metrics $DM --clock-control none --import-source on -k regex:k_step -s 8 -c 1 -f -o $out/${tag}_kstep_inloop \
    python scripts/prof_k1.py solve > $out/${tag}_kstep_inloop.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_line.json 2> $out/${tag}_bench_reference.err
timeout 900 python bench.py > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err
tail -c 600 $out/${tag}_bench_line.json
