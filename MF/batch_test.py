"""Import-time argument parsing and data load: `from batch_test import *` must hand train_new_api.py the names
args / data / Ks / BATCH_SIZE / ITEM_NUM / USER_NUM, exactly like the reference's MF/batch_test.py:1-20 does."""
import ast

import load_data
import parse

args = parse.parse_args()

# models trained on (user, item, stage) triples -- PD / PDA / BPR(t)-pop -- read train_with_time.txt (Data2);
# BPRMF / PDG read train.txt (Data)
_STAGED_MODELS = frozenset({"s_condition", "sg_condition", "temp_pop", "us_condition"})
data = (load_data.Data2 if args.train in _STAGED_MODELS else load_data.Data)(args)

Ks = list(ast.literal_eval(args.Ks))      # "--Ks [20,50]"
BATCH_SIZE, ITEM_NUM, USER_NUM = args.batch_size, data.n_items, data.n_users
