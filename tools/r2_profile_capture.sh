set -x
B="python bench.py --batch 32 --max-batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-leg --no-check --sharded-api off"
$B > gpurun_out/r2_plain_b32.json 2> gpurun_out/r2_plain_b32.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 900 -c 900 --csv --log-file gpurun_out/r2_launches_b32.csv $B > gpurun_out/r2_ncu_b32.log 2>&1
AB_WARM=0 AB_REPS=1 AB_B=16 python tools/attn_bench.py 3 > gpurun_out/r2_plain_attn.log 2>&1 && \
AB_WARM=0 AB_REPS=1 AB_B=16 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -c 1 -f -o gpurun_out/r2_prof_attn python tools/attn_bench.py 3 > gpurun_out/r2_ncu_attn.log 2>&1
AB_WARM=0 AB_REPS=1 AB_B=16 python tools/attn_bench.py 2 > gpurun_out/r2_plain_attn_x3.log 2>&1 && \
AB_WARM=0 AB_REPS=1 AB_B=16 ncu --set full --clock-control none --import-source on -k regex:attention_tc_x3 -c 1 -f -o gpurun_out/r2_prof_attn_x3 python tools/attn_bench.py 2 > gpurun_out/r2_ncu_attn_x3.log 2>&1
ls -la gpurun_out/r2_prof* gpurun_out/r2_launches_b32.csv
