set -x
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_v5.json 2> gpurun_out/r2_bench_reference_v5.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_v5.json 2> gpurun_out/r2_bench_v5.err
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_v5.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r2_smoke_v5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lob_replay -s 2 -c 1 -o gpurun_out/prof_replay_r2v5 python tools/kbench.py --iters 1 --replay-msgs 38400 > gpurun_out/ncu_replay_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lob_step -s 6 -c 2 -o gpurun_out/prof_step_deep_r2v5 python tools/kbench.py --skip-replay --iters 1 --config hetero_deep_book > gpurun_out/ncu_step_deep_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lob_step_prep|lob_step_scan|lob_agents_finish|lob_step_reset_done" -s 12 -c 4 -o gpurun_out/prof_step_piped_r2v5 python tools/kbench.py --skip-replay --iters 1 > gpurun_out/ncu_step_r2v5.log 2>&1
python tools/rollout_time.py > gpurun_out/r2_rollout_time_v5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu --env-inner 4 > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out | tail -20
