mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench8 rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/r02_bench_n8.json")); print("ms/step", round(j["ms_per_step"],3), "value", round(j["value"]/1e6,1), "e2e", round(j["e2e"]["value"]/1e6,1), "step_kernel", round(j["kernels"]["bpr_step"]["ms_per_launch"],3), "adam", round(j["kernels"]["adam_apply"]["ms_per_step"],3), j["exchange"], j["parity"], "eval", j["eval"]["value"]/1e12, j["clocks"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --config kwai --gpus 8 > gpurun_out/r02_config_kwai_n8.json 2> gpurun_out/r02_config_kwai_n8.err; echo "kwai8 rc=$?"; cat gpurun_out/r02_config_kwai_n8.json | cut -c1-900
