#!/bin/bash
# Trimmed refresh of the round-2 artefacts after a kernel-policy change (about 10 GPU-minutes):
#   gpurun --timeout 1500 -- 'bash scripts/r2_profile_quick.sh'
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --profile-out gpurun_out/r2_kernel_table.json > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 200 gpurun_out/r2_bench_1gpu.json
timeout 300 python bench.py --batch 512 --size 112 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_cfg3_112_b512.json 2>&1
timeout 300 python bench.py --batch 256 --height 192 --width 256 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_cfg4_192x256.json 2>&1
timeout 300 python bench.py --batch 256 --height 256 --width 192 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2_bench_cfg4_256x192.json 2>&1
timeout 600 python bench.py --global-batch 2048 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_cfg2_strong_2048_1gpu.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k 'regex:_k$' --csv --log-file gpurun_out/ncu_raw.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --dump-ops gpurun_out/ops.json > gpurun_out/ncu_bench.log 2>&1
python scripts/ncu_summarize.py gpurun_out/ncu_raw.csv gpurun_out/ops.json r2 gpurun_out > gpurun_out/ncu_summarize.log 2>&1 || tail -3 gpurun_out/ncu_summarize.log
rm -f gpurun_out/ncu_raw.csv
timeout 900 ncu --set full --clock-control none -k 'regex:conv_tc_k|pw_stream_k|pw_wgrad_stream_k|dw_mma|dws_|c3_|stem_fwd_mma|pw_bwd_fused|pw_proj_bwd|dw_tile_k' \
    --launch-skip 700 --launch-count 80 -o /tmp/r2_ncu_step_gemm_dw -f \
    python bench.py --steps 3 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python scripts/ncu_summary.py /tmp/r2_ncu_step_gemm_dw.ncu-rep gpurun_out/r2_ncu_gemm_dw_summary.json > /dev/null 2>&1
timeout 300 python scripts/exp_dw_small.py parity time bwd > gpurun_out/r2_exp_dw_small.txt 2>&1; tail -3 gpurun_out/r2_exp_dw_small.txt | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -1 gpurun_out/r2_smoke.txt
find gpurun_out -size +20M -delete
echo done
