for se in 8 12 16 24; do
PDA_TC_SE=$se timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ev_se$se.json 2>/dev/null
python - <<PY
import json
j=json.load(open("gpurun_out/ev_se$se.json"))["eval"]
for t in ("trained","fitted"):
    e=j[t]; print("se=$se", t, "ms", round(e["kernel_ms"],3), "frac_burst", round(e["roofline"]["frac"],4), "passB", round(e["sweep_pass_b"]["ms"],3), "passA", round(e["sweep_pass_a"]["ms"],3), "other", round(e["other_kernels_ms"],3), "cand/row", round(e["candidates_per_row"],1), "fallback", e["rows_exact_fallback"], "stride", e["filter_stats"]["tile_stride"])
PY
done
