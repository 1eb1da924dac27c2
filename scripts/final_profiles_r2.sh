# Round-2 measurement set (one B200).  Plain runs first; every ncu pass only after its own command has exited 0 without ncu.
set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
python scripts/bench_linear.py > gpurun_out/r2f_bench_linear.txt 2>&1
MSMP_LINEAR_TS=0 python scripts/bench_linear.py > gpurun_out/r2f_bench_linear_tma.txt 2>&1
MSMP_PRECISION=bf16 python scripts/bench_linear.py > gpurun_out/r2f_bench_linear_reduced.txt 2>&1
python scripts/bench_wgrad.py > gpurun_out/r2f_bench_wgrad.jsonl 2>&1
# serialised step: per-launch time and DRAM bytes
python scripts/ncu_step.py c4 8 > gpurun_out/r2f_plain_step.log 2>&1 && \
timeout 800 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches_c4_step.csv python scripts/ncu_step.py c4 8 > gpurun_out/r2f_ncu_step.log 2>&1; echo "launch list rc=$?"
# full captures of the node GEMMs (one launch per shape) and the weight-gradient kernels
BENCH_REPS=1 python scripts/bench_linear.py > /dev/null 2>&1 && \
BENCH_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_linear_ts -o gpurun_out/r2f_linear_ts -f python scripts/bench_linear.py > gpurun_out/r2f_ncu_linear.log 2>&1; echo "ncu linear rc=$?"
python scripts/ncu_wgrad.py > /dev/null 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wgrad -o gpurun_out/r2f_wgrad -f python scripts/ncu_wgrad.py > gpurun_out/r2f_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
ls -la gpurun_out/ | grep r2f
