set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1f_bench_ref.json 2> gpurun_out/r1f_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/r1f_plain_eager.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1400 --csv --log-file gpurun_out/r1f_launches_bench_eager.csv python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > gpurun_out/r1f_ncu_launch.log 2>&1; echo "launch list rc=$?"
python scripts/ncu_kernels.py > gpurun_out/r1f_plain_k.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lem|k_wgrad_tc|k_linear_tc|k_edge_ws|k_segment_reduce" -c 24 -o gpurun_out/r1f_c2_kernels -f python scripts/ncu_kernels.py > gpurun_out/r1f_ncu_k.log 2>&1; echo "ncu c2 rc=$?"
REPS=2 python scripts/edge_ticks.py > gpurun_out/r1f_plain_edge.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_edge_ws -s 2 -c 2 -o gpurun_out/r1f_edge_ws_large -f env REPS=2 python scripts/edge_ticks.py > gpurun_out/r1f_ncu_edge.log 2>&1; echo "ncu edge rc=$?"

python scripts/timeline.py > gpurun_out/r1f_timeline.txt 2>&1; rm -f gpurun_out/trace.json
python scripts/bench_configs.py > gpurun_out/r1f_bench_configs.jsonl 2>&1
python scripts/bench_large.py > gpurun_out/r1f_bench_large.jsonl 2>&1
python scripts/edge_sweep.py > gpurun_out/r1f_edge_sweep.txt 2>&1
ls -la gpurun_out/ | tail -25
