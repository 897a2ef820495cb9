for cfg in "1024 512" "4096 2048" "8192 2048"; do set -- $cfg; PDA_TC_CAP=$1 PDA_TC_RC=$2 timeout 300 python bench.py --config douban_pda_eval 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read());
for w in ('valid','test'):
    e=j[w]; print('cap $1 rc $2', w, 'kernel_ms', round(e['kernel_ms'],3), 'fallback', e['filter_stats']['rows_exact_fallback'], 'overflow', e['filter_stats']['rows_overflow'], 'cand', e['filter_stats']['candidates'], 'max', e['filter_stats']['max_candidates_row'], 'ids_equal', e['ids_equal_oracle'])"; done
