# Round-2 evidence run on ONE B200 (gpurun): GPU test log, ncu launch list + full captures, bench lines of every config.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests.log; tail -3 gpurun_out/r02_gpu_tests.log)
# 1. launch list of the bench command (shares per kernel; 3 + 5 steps, eval of 16384 users)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pda|bpr_|adam_|sample_kernel|tc_|recommend_|finish_|segsum|xavier" -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --eval-users 16384 > gpurun_out/r02_launch_bench.json 2> gpurun_out/r02_launch_bench.err
# 2. full capture of the training kernels at steady state (launches well after the pre-aging)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bpr_step_pipe_kernel|adam_dense_kernel|sample_kernel" -s 120 -c 3 -o gpurun_out/r02_train python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --no-eval > /dev/null 2> gpurun_out/r02_train_ncu.err
# 3. full capture of the eval kernels (one 16384-user block, fitted tables)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_" -s 14 -c 14 -o gpurun_out/r02_eval python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --eval-users 16384 --eval-tables fitted > /dev/null 2> gpurun_out/r02_eval_ncu.err
# 4. the bench lines
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 8 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
for c in douban_pd douban_pda_eval kwai; do timeout 600 python bench.py --config $c > gpurun_out/r02_config_$c.json 2> gpurun_out/r02_config_$c.err; echo "$c rc=$?"; done
ls -la gpurun_out | tail -20
