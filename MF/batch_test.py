"""Import-time argument parsing and data load, like the reference's MF/batch_test.py:1-20."""
from parse import parse_args
from load_data import Data, Data2

args = parse_args()

if args.train in ('s_condition', 'sg_condition', 'temp_pop', 'us_condition'):
    data = Data2(args)      # PD / PDA / BPR(t)-pop: interactions with stage labels
else:
    data = Data(args)       # BPRMF / PDG

Ks = eval(args.Ks)
BATCH_SIZE = args.batch_size
ITEM_NUM = data.n_items
USER_NUM = data.n_users
