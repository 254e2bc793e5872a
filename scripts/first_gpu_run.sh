#!/bin/bash
# One gpurun call that refreshes every measured artefact for the shipped defaults and times the opt-in kernels:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/first_gpu_run.sh'
# Everything lands in gpurun_out/ (copy what should be judged into profiles/).  ~10 GPU-minutes on one B200.
set -u
mkdir -p gpurun_out
echo "== gpu tests ==";      timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke ==";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
echo "== bench (1 GPU) ==";  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --profile-out gpurun_out/kernel_table.json > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 600 gpurun_out/bench_1gpu.json
echo "== 1x1 streaming vs tcgen05 =="
timeout 300 python scripts/exp_stream.py pw_big stem_big net > gpurun_out/exp_stream.log 2>&1; tail -2 gpurun_out/exp_stream.log | cut -c1-600
echo "== ncu launch list of one timed step (shares + DRAM bytes; never a bench value) =="
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k 'regex:_k$' --csv --log-file gpurun_out/ncu_raw.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --dump-ops gpurun_out/ops.json > gpurun_out/ncu_bench.log 2>&1
# -> gpurun_out/r2_ncu_step_summary.json + r2_ncu_launch_list.csv (copy into profiles/); the assert inside checks that
# the launch list is a whole number of steps of the dumped op sequence
python scripts/ncu_summarize.py gpurun_out/ncu_raw.csv gpurun_out/ops.json r2 gpurun_out > gpurun_out/ncu_summarize.log 2>&1 || tail -3 gpurun_out/ncu_summarize.log
rm -f gpurun_out/ncu_raw.csv        # tens of MB; the summary and the per-launch list are what is kept
echo done
