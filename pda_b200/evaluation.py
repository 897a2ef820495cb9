"""Evaluation driver: all eval users x all items -> top-50 -> Recall / Precision / NDCG / Hit @Ks.

Mirrors the `evaluation` class of the reference (MF/train_new_api.py:700-794): set_evaluate_obj_pre builds
the user batches and the train-item mask, eval() runs the recommender per batch and averages
get_performance (MF/used_metric.py:69-80) over the eval users.  Here the mask is the train CSR already on the
device (it holds exactly the (row, train item) pairs of :730-733), the top-K ids never go through a Pool of
python workers, and the metrics are reduced by the library (pda_metrics_host).  Batches default to ALL eval
users in one call (the kernels tile internally); `batch_size` users per call are used only where the
reference's batching is visible in the result (BPR(t)-pop: the user-bias factor comes from the first user of
each 2048-user batch).
"""
from __future__ import annotations

import numpy as np


class evaluation:
    def __init__(self, data, Ks, batch_size=2048):
        self.data, self.Ks = data, list(Ks)
        self.batch_size = batch_size
        self.testing_popularity = None
        self.set_evaluate_obj()

    def set_evaluate_obj(self, eval_who='test'):
        self.eval_who = eval_who

    def set_testing_popularity(self, popularity):
        self.testing_popularity = popularity

    def set_evaluate_obj_pre(self, eval_who='test'):
        d = self.data
        self.eval_who = eval_who
        self.eval_user_list = d.test_user_list if eval_who == 'test' else d.valid_user_list
        self.truth_indptr = d.test_indptr if eval_who == 'test' else d.valid_indptr
        self.truth_items = d.test_items if eval_who == 'test' else d.valid_items
        self.all_users = np.asarray(self.eval_user_list.keys(), dtype=np.int32)     # file order, like dict.keys()
        self.tot_user = len(self.all_users)
        bs = self.batch_size
        self.list_batch_user = [self.all_users[i:i + bs] for i in range(0, self.tot_user, bs)]

    def eval(self, model, sess=None, rec_type='main_branch'):
        """-> {'precision','recall','ndcg','hit_ratio'}: np.float64 arrays over Ks (means over eval users)."""
        pos_pop = None if self.testing_popularity is None else np.asarray(self.testing_popularity)
        per_batch = getattr(model, "needs_reference_eval_batches", False)
        batches = self.list_batch_user if per_batch else [self.all_users]
        result = {k: np.zeros(len(self.Ks)) for k in ('precision', 'recall', 'ndcg', 'hit_ratio')}
        for batch_user in batches:
            if len(batch_user) == 0:
                continue
            ids = model.do_recommendation(sess, batch_user, None, rec_type, pos_pop=pos_pop)
            s = model.metrics_sum(ids, batch_user, self.truth_indptr, self.truth_items, self.Ks)
            for k in result:
                result[k] += s[k] / self.tot_user
        return result
