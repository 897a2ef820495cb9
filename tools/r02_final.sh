# Final evidence run of round 2 on ONE B200: GPU tests, smoke, launch list + full capture of the training kernels, bench line.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02f_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_gpu_tests.log; tail -3 gpurun_out/r02f_gpu_tests.log)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pda|bpr_|adam_|sample_kernel|tc_|recommend_|finish_|segsum|xavier|item_count" -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --eval-users 16384 > gpurun_out/r02f_launch_bench.json 2> gpurun_out/r02f_launch_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bpr_step_pipe_kernel|adam_dense_kernel|sample_kernel" -s 120 -c 3 -o gpurun_out/r02f_train python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --no-eval > /dev/null 2> gpurun_out/r02f_train_ncu.err
timeout 600 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02f_bench_n1.json
