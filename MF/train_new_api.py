"""python -u MF/train_new_api.py --dataset douban --train s_condition --test s_condition --pop_exp 0.22 ...

Drop-in for the reference's MF/train_new_api.py (same flags, stdout formats and checkpoint schedule); the
session, sampler, model graph, optimizer, scorer and metric code behind it are libpda_b200.so on a B200.
The driver itself lives in pda_b200/driver.py.
"""
import os
import signal
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from batch_test import *  # noqa: E402,F401,F403  (parses the flags and loads the data, like the reference)
from pda_b200.driver import DatasetApi_Model, early_stop, main  # noqa: E402,F401
from pda_b200.evaluation import evaluation  # noqa: E402,F401


def term(sig_num, addtion):
    # the reference SIGKILLs its process group so forked sampler workers die with it (:48-51); no workers here
    sys.exit(1)


signal.signal(signal.SIGTERM, term)

if __name__ == '__main__':
    main(args, data, Ks)  # noqa: F405
