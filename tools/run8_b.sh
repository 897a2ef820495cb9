mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/dpx_bench.py 2>&1 | grep "PDA_DPX" | tee gpurun_out/dpx8.jsonl
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 60 --warmup 5 --no-eval --no-cpu > gpurun_out/b8_$tag.json 2> gpurun_out/b8_$tag.err; echo "$tag rc=$?"; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/b8_$tag.json")); print("$tag", "ms/step", round(j["ms_per_step"],3), "value", round(j["value"]/1e6,1), "e2e", round(j["e2e"]["value"]/1e6,1), "step_kernel", round(j["kernels"]["bpr_step"]["ms_per_launch"],3), "adam", round(j["kernels"]["adam_apply"]["ms_per_step"],3), j["exchange"], j["parity"]["all_true"] if j.get("parity") else None)
except Exception as e: print("$tag failed", e)
PY
}
run nvls PDA_DP_EXCHANGE=auto
run scatter2 PDA_DP_EXCHANGE=scatter
