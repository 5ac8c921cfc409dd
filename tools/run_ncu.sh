#!/bin/bash
# The ncu passes whose summaries are committed under profiles/ (B200_PROFILING.md recipe).  Each workload is
# first run WITHOUT ncu and must exit 0.  Outputs: gpurun_out/ncu_r2_*.{csv,log,ncu-rep}
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. DRAM traffic of every fused-attention launch of one bench job (one pass: two dram counters + duration)
python tools/ncu_jobs.py attn_job c4 > gpurun_out/ncu_r2_attn_job_plain.json 2> gpurun_out/ncu_r2_attn_job_plain.err || exit 1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:decode_attention --csv \
     --log-file gpurun_out/ncu_r2_attn_traffic_c4.csv python tools/ncu_jobs.py attn_job c4 > gpurun_out/ncu_r2_attn_job_ncu.json 2> gpurun_out/ncu_r2_attn_job_ncu.err
echo "attn traffic rc=$?"
# 2. --set full of the prefill-sized merged GEMM launch
python tools/ncu_jobs.py prefill > gpurun_out/ncu_r2_prefill_plain.json 2> gpurun_out/ncu_r2_prefill_plain.err || exit 1
$NCU --set full --import-source on -k regex:gemm_tf32x3 -o gpurun_out/ncu_r2_prefill -f python tools/ncu_jobs.py prefill > gpurun_out/ncu_r2_prefill_ncu.log 2>&1
echo "prefill full rc=$?"
ncu -i gpurun_out/ncu_r2_prefill.ncu-rep --page raw --csv > gpurun_out/ncu_r2_prefill_raw.csv 2>/dev/null
# 3. --set full of one fused-attention launch (warp-per-position kernel)
python tools/ncu_jobs.py attn_one > gpurun_out/ncu_r2_attn_one_plain.json 2> gpurun_out/ncu_r2_attn_one_plain.err || exit 1
$NCU --set full --import-source on -k regex:decode_attention --launch-skip 2 --launch-count 1 -o gpurun_out/ncu_r2_attn_one -f python tools/ncu_jobs.py attn_one > gpurun_out/ncu_r2_attn_one_ncu.log 2>&1
echo "attn full rc=$?"
ncu -i gpurun_out/ncu_r2_attn_one.ncu-rep --page raw --csv > gpurun_out/ncu_r2_attn_one_raw.csv 2>/dev/null
# 4. launch list of the bench command itself (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_r2_bench_plain.json 2> gpurun_out/ncu_r2_bench_plain.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/ncu_r2_launches_c4.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_r2_bench_ncu.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/ | grep ncu_r2
# 5. emb_dim 4096 prefill launch: DRAM traffic and tensor-pipe activity with the grouped item order and with the plain one
python tools/ncu_jobs.py prefill_d4096 > gpurun_out/ncu_r2_prefill4096_plain.json 2> gpurun_out/ncu_r2_prefill4096_plain.err || exit 1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
$NCU --metrics $M -k regex:gemm_tf32x3_pair --csv --log-file gpurun_out/ncu_r2_prefill4096_grouped.csv python tools/ncu_jobs.py prefill_d4096 > /dev/null 2>&1
MLI_TC_KV_GROUP=32 $NCU --metrics $M -k regex:gemm_tf32x3_pair --csv --log-file gpurun_out/ncu_r2_prefill4096_plainorder.csv python tools/ncu_jobs.py prefill_d4096 > /dev/null 2>&1
echo "prefill d4096 rc=$?"
