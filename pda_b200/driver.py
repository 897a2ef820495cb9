"""Training / evaluation driver behind MF/train_new_api.py (the reference's CLI, kept as the entry point).

Mirrors MF/train_new_api.py of the reference: DatasetApi_Model (:536-696), early_stop (:910-926) and the main
program (:930-1340) -- same flags, same model selection by --train, same epoch / evaluation / early-stop /
checkpoint schedule and the same stdout line formats (:1110, :1119-1123, :1135, ...).  What differs is below the
API: there is no TF session -- the sampler, the fused BPR step, the TF1-semantics Adam sweep, the scoring +
top-50 and the metrics all run in libpda_b200.so on the GPU, and an epoch is enqueued as n_batch x
(sample, step, adam) without a host round trip per step.
"""
from __future__ import annotations

import os
import sys
from time import time

import numpy as np

from . import popularity as popmod
from .evaluation import evaluation
from .model import PDAModel, TOPK_MAX

SAMPLER_SEED = 2020   # random.seed(2020); np.random.seed(2020)   (train_new_api.py:934-935)
INIT_SEED = 2021      # tf.set_random_seed(2021)                   (:936)


class DatasetApi_Model:
    """train_new_api.py:536-696.  `sess` arguments are accepted and ignored (there is no session)."""

    def __init__(self, args, data_config, test_batch, data, device=0):
        if args.train in ('s_condition', 'condition', 'temp_pop'):
            self.input_type = "with_pop"
            print("dataset api with pop or temp")
            if args.train == 'temp_pop':
                self.input_type = 'with_temp'
        else:
            self.input_type = 'without_pop'
            print("dataset api without pop")
        if args.train not in ('normal', 's_condition', 'condition', 'temp_pop'):
            raise NotImplementedError("not implement this model: " + args.train)
        self.args, self.data = args, data
        self.n_items = data_config['n_items']
        kw = {}
        if args.train == 'temp_pop':
            kw["temp_num"] = data_config['temp_num']
        self.Recommender = PDAModel(data_config['n_users'], data_config['n_items'], args.embed_size, train=args.train,
                                    batch_size=args.batch_size, lr=args.lr, regs=args.regs, device=device,
                                    seed=INIT_SEED, **kw)
        adam_mode = os.environ.get("PDA_ADAM_MODE", "")     # "dense" | "lazy": same results bit for bit
        if adam_mode and args.train != 'temp_pop':
            self.Recommender.set_adam_mode(adam_mode)
        self.Recommender.set_train_csr(data.train_indptr, data.train_items, data.train_times,
                                       unique_times=data.unique_times or None)
        self.needs_reference_eval_batches = args.train == 'temp_pop'
        self.testing_model_type, self.testing_popularity, self.sess = 'o', None, None
        self._epoch = -1
        self._pop_set = False

    # -- sampler side: data.add_expo_popularity(...) of the reference feeds the generators; here the table
    #    goes to the device once, right before the first epoch
    def _ensure_pop(self):
        if self._pop_set:
            return
        if self.args.train in ('s_condition', 'condition'):
            if self.data.expo_popularity is None:
                raise RuntimeError("data.add_expo_popularity(...) must be called before training PD / PDG")
            self.Recommender.set_train_pop(np.asarray(self.data.expo_popularity, dtype=np.float64).astype(np.float32))
        self._pop_set = True

    def switch_to_training_or_reinitsampler(self, sess=None):
        self._epoch += 1     # a new epoch = a new stream of sampled batches

    def train_epoch(self, n_batch):
        """n_batch x sess.run([opt, loss, mf_loss, reg_loss]) -> epoch means (loss, mf_loss, reg_loss)."""
        self._ensure_pop()
        r = self.Recommender
        r.read_loss_sums(reset=True)
        r.train_sampled(SAMPLER_SEED, self._epoch, 0, n_batch, self.args.batch_size)
        s = r.read_loss_sums(reset=True)
        return s[0] / n_batch, s[1] / n_batch, s[2] / n_batch

    def train_one_batch(self, step):
        """one sess.run([opt, loss, mf_loss, reg_loss]) of the current epoch."""
        self._ensure_pop()
        r = self.Recommender
        r.train_sampled(SAMPLER_SEED, self._epoch, step, 1, self.args.batch_size)
        return r.read_loss()

    # -- inference side
    def do_recommendation(self, sess, batch_users, items, rec_type, pos_pop=None, sparse_cliked_matrix=None):
        if rec_type not in ('main_branch', 'main_with_pop', 'condition'):
            raise NotImplementedError("we have only implement recommendation method: main main+pop condition")
        if rec_type == 'condition' and self.args.train not in ('s_condition', 'condition'):
            raise NotImplementedError("condition_ratings exist only for the PD / PDG graph")
        col_bias = None
        if self.input_type == 'with_temp':
            col_bias = self.Recommender.temp_item_bias_for_eval(int(batch_users[0]))
        return self.Recommender.do_recommendation(batch_users, items, rec_type, pos_pop=pos_pop, K=TOPK_MAX,
                                                  col_bias=col_bias)

    def metrics_sum(self, ids, eval_users, truth_indptr, truth_items, Ks):
        return self.Recommender.metrics_sum(ids, eval_users, truth_indptr, truth_items, Ks)

    def testing(self, sess, batch_users, items, model_type, pos_pop=None):
        return self.Recommender.testing(batch_users, items, model_type, pos_pop=pos_pop)

    def set_testing_way(self, model_type, popularity_exp):
        self.testing_model_type, self.testing_popularity = model_type, popularity_exp
        self.Recommender.set_testing_way(model_type, popularity_exp)

    def set_sess(self, sess):
        self.sess = sess

    def predict(self, user_batch, item_batch=None):
        return self.Recommender.predict(user_batch, item_batch)

    # -- tf.train.Saver stand-in: one .npz per checkpoint, TF variable names as keys
    def save(self, path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez(path + ".npz", **self.Recommender.state_dict())

    def restore(self, path):
        self.Recommender.load_state_dict(dict(np.load(path + ".npz")))


def early_stop(hr, ndcg, recall, precision, cur_epoch, config, stopping_step, flag_step=10):
    """train_new_api.py:910-926: model selection on recall@Ks[0] (ties count as improvement)."""
    if recall >= config['best_recall']:
        stopping_step = 0
        config.update(best_hr=hr, best_ndcg=ndcg, best_recall=recall, best_pre=precision, best_epoch=cur_epoch)
    else:
        stopping_step += 1
    should_stop = stopping_step >= flag_step
    if should_stop:
        print("Early stopping is trigger")
    return config, stopping_step, should_stop


def print_result_f(ret):
    print('||---------------------------------------------- recall=[%.5f, %.5f], precision=[%.5f, %.5f], '
          'hit=[%.5f, %.5f], ndcg=[%.5f, %.5f]' % (ret['recall'][0], ret['recall'][-1], ret['precision'][0],
                                                    ret['precision'][-1], ret['hit_ratio'][0], ret['hit_ratio'][-1],
                                                    ret['ndcg'][0], ret['ndcg'][-1]))


def _gamma_tilde_search(evaluation_model, model, last_pop_ori, best_ret):
    """BPRMF-A: line search of gamma~ over 0.04, 0.06, ... until 5 non-improvements (train_new_api.py:1170-1187)."""
    best_expo, not_incre, expo = 0, 0, 0.04
    while True:
        evaluation_model.set_testing_popularity(np.power(last_pop_ori, expo))
        ret_k = evaluation_model.eval(model, None, rec_type='main_with_pop')
        if ret_k['recall'][0] < best_ret['recall'][0]:
            not_incre += 1
            if not_incre > 4:
                break
        else:
            not_incre, best_ret, best_expo = 0, ret_k, expo
        print("expo: {:.2f} best expo:{:.2f}".format(expo, best_expo))
        print_result_f(ret_k)
        expo += 0.02
    return best_ret, best_expo


def main(args, data, Ks):
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", str(args.cuda))
    config = dict(n_users=data.n_users, n_items=data.n_items)
    popularity_exp = args.pop_exp
    print("----- popularity_exp : ", popularity_exp)
    test_batch_size = min(1024, args.batch_size)

    pop_item_all = popmod.load_popularity(args)
    last_stage_popualarity, linear_predict_popularity = popmod.eval_popularities(pop_item_all, popularity_exp)

    if args.model == 'mf' and args.train == 'normal':
        args.saveID += "pop_exp-{:.2f}".format(popularity_exp)
        print("normal MF... ")
        model = DatasetApi_Model(args, config, test_batch_size, data)
        last_stage_popualarity_ori, linear_predict_popularity_ori = popmod.bprmf_a_popularities(
            pop_item_all, linear_predict_popularity)
    elif args.model == 'mf' and args.train == 'condition':
        args.saveID += "pop_exp-{:.2f} (gamma)".format(popularity_exp)
        print("PD-G & PDG-A based on MF... ")
        print("compute tot popularity...")
        popularity_matrix = np.power(popmod.get_dataset_tot_popularity(data), popularity_exp)
        print("After power,popularity information-- mean:{},max:{},min:{}".format(
            popularity_matrix.mean(), popularity_matrix.max(), popularity_matrix.min()))
        model = DatasetApi_Model(args, config, test_batch_size, data)
        data.add_expo_popularity(popularity_matrix)
    elif args.model == 'mf' and args.train == 's_condition':
        print('-------    running PD & PDA model  ----------------')
        args.saveID += "pop_exp-{:.2f} (gamma)".format(popularity_exp)
        print("save_ID", args.saveID)
        popularity_matrix = np.power(popmod.get_popularity_from_load(pop_item_all), popularity_exp)
        print("------ popularity information after powed  ------")
        print("   each stage mean:", popularity_matrix.mean(axis=0))
        print("   each stage max:", popularity_matrix.max(axis=0))
        print("   each stage min:", popularity_matrix.min(axis=0))
        model = DatasetApi_Model(args, config, test_batch_size, data)
        data.add_expo_popularity(popularity_matrix)
    elif args.model == 'mf' and args.train == 'temp_pop':
        print('-------    running temproal pop MF  ----------------')
        config['temp_num'] = pop_item_all.shape[1] - 1
        args.saveID += "temp_pop"
        print("save_ID", args.saveID)
        model = DatasetApi_Model(args, config, test_batch_size, data)
        data.add_expo_popularity(None)
    else:
        raise NotImplementedError("do not implement this method")

    args.wd = args.regs
    evaluation_model = evaluation(data, Ks)
    if args.valid_set == "test":
        evaluation_model.set_evaluate_obj_pre('test')
        print("valid in test set")
    elif args.valid_set == "valid":
        print("valid in valid set")
        evaluation_model.set_evaluate_obj_pre('valid')
    else:
        print('evaluate type error.')
        sys.exit()
    print("args info:", args)
    print("top K:", Ks)

    best_pop_expo_normal = 0
    config.update(best_hr=0, best_ndcg=0, best_recall=0, best_pre=0, best_epoch=0, best_c_hr=0, best_c_epoch=0, best_c=0.0)
    config_main = config.copy()
    stopping_step_main = stopping_step = 0
    n_batch = data.n_train // args.batch_size + 1
    t1 = time()
    print("batch_num:", n_batch, 'waiting sampling...')
    save_ckpt_dir = args.save_dir + '{}_{}_checkpoint/wd_{}_lr_{}_a_{}_{}_train_{}/'.format(
        args.model, args.dataset, args.wd, args.lr, args.alpha, args.saveID, args.train)
    for epoch in range(args.epoch):
        model.switch_to_training_or_reinitsampler(None)
        loss, mf_loss, reg_loss = model.train_epoch(n_batch)
        if np.isnan(loss):
            print('ERROR: loss is nan.')
            sys.exit()
        if epoch % args.log_interval != 0:
            if args.verbose > 0 and epoch % args.verbose == 0:
                print('Epoch %d [%.1fs]: train==[%.5f=%.5f + %.5f]' % (epoch, time() - t1, loss, mf_loss, reg_loss))
            t1 = time()
            continue

        perf_str = 'Epoch %d [%.1fs]: train==[%.5f=%.5f + %.5f]' % (epoch, time() - t1, loss, mf_loss, reg_loss)
        if args.test in ('condition', 's_condition'):
            print("do not consider popularity (PD or PDG) ... ")
            print(perf_str)
            evaluation_model.set_testing_popularity(None)
            ret_main = evaluation_model.eval(model, None, rec_type='main_branch')
            print_result_f(ret_main)
            print("injecting last stage popularity.... ")
            ttt1 = time()
            evaluation_model.set_testing_popularity(last_stage_popualarity)
            ret1 = evaluation_model.eval(model, None, rec_type='condition')
            print("||------------PDA/PDGA injecting last stage popularity testing : time: ", int(time() - ttt1))
            print_result_f(ret1)
            ttt1 = time()
            evaluation_model.set_testing_popularity(linear_predict_popularity)
            ret2 = evaluation_model.eval(model, None, rec_type='condition')
            print("||------------PDA/PDGA injecting linear predicted popularity testing : time: ", int(time() - ttt1))
            print_result_f(ret2)
            ret = ret1
        elif args.test == "normal":
            print(perf_str)
            ttt1 = time()
            evaluation_model.set_testing_popularity(None)
            ret_main = evaluation_model.eval(model, None, rec_type='main_branch')
            print("test: time:", time() - ttt1)
            print_result_f(ret_main)
            best_ret, best_expo = _gamma_tilde_search(evaluation_model, model, last_stage_popualarity_ori, ret_main)
            if best_ret['recall'][0] >= config['best_recall']:
                best_pop_expo_normal = best_expo
            ret = best_ret
        elif args.test == 'temp_pop':
            print(perf_str)
            ttt1 = time()
            evaluation_model.set_testing_popularity(None)
            ret_main = evaluation_model.eval(model, None, rec_type='main_branch')
            print("test: time:", time() - ttt1)
            print_result_f(ret_main)
            ret = ret_main
        else:
            raise NotImplementedError("not implement this test method:" + args.test)

        stop_flag_step = 100 // args.log_interval
        config, stopping_step, should_stop = early_stop(ret['hit_ratio'][0], ret['ndcg'][0], ret['recall'][0],
                                                        ret['precision'][0], epoch, config, stopping_step,
                                                        flag_step=stop_flag_step)
        config_main, stopping_step_main, should_stop_main = early_stop(
            ret_main['hit_ratio'][0], ret_main['ndcg'][0], ret_main['recall'][0], ret_main['precision'][0], epoch,
            config_main, stopping_step_main, flag_step=stop_flag_step)
        if epoch == config['best_epoch']:
            model.save(save_ckpt_dir + "best_ckpt.ckpt")
        if epoch == config_main['best_epoch']:
            model.save(save_ckpt_dir + "best_main_ckpt.ckpt")
        if args.save_flag == 1 and (epoch + 1) % 50 == 0:
            model.save(save_ckpt_dir + "{}_ckpt.ckpt".format(epoch))
        if should_stop and args.early_stop == 1 and should_stop_main:
            print("{} dataset best epoch{}: hr:{} ndcg:{} recall:{} precision:{}".format(
                args.dataset, config['best_epoch'], config['best_hr'], config['best_ndcg'], config['best_recall'],
                config['best_pre']))
            print("{} dataset best main epoch{}: hr:{} ndcg:{} recall:{} precision:{}".format(
                args.dataset, config_main['best_epoch'], config_main['best_hr'], config_main['best_ndcg'],
                config_main['best_recall'], config_main['best_pre']))
            if args.save_flag == 1:
                with open(save_ckpt_dir + '/best_epoch.txt', 'w') as f:
                    print(config['best_epoch'], file=f)
            break
        t1 = time()

    # ---- performance on the testing set at the best epoch (train_new_api.py:1248-1335) ----
    print("best epoch", config['best_epoch'])
    model.restore(save_ckpt_dir + "best_ckpt.ckpt")
    print("validation result in best epoch")
    evaluation_model.set_testing_popularity(None)
    ret = evaluation_model.eval(model, None, rec_type='main_branch')
    print("---- result without pop:")
    print_result_f(ret)
    print("|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||")
    print("|| ---------------- testing testset in the best epoch:  ... ")
    evaluation_model.set_evaluate_obj_pre('test')
    final = {}
    if args.test in ('s_condition', 'condition'):   # the reference forgets 'condition' here and raises (:1330)
        evaluation_model.set_testing_popularity(None)
        ret = evaluation_model.eval(model, None, rec_type='main_branch')
        print("---- PD/PDG result without pop at the model select by PDA/PDG-A:")
        print_result_f(ret)
        final["main_branch"] = ret
        evaluation_model.set_testing_popularity(last_stage_popualarity)
        ret = evaluation_model.eval(model, None, rec_type='condition')
        print("---- PDA/PDG-A injecting last stage pop:\n", ret)
        print_result_f(ret)
        final["condition_last"] = ret
        evaluation_model.set_testing_popularity(linear_predict_popularity)
        ret = evaluation_model.eval(model, None, rec_type='condition')
        print("---- result with linear pop:\n", ret)
        print_result_f(ret)
        final["condition_linear"] = ret
    elif args.test == "normal":
        evaluation_model.set_testing_popularity(None)
        ret = evaluation_model.eval(model, None, rec_type='main_branch')
        print("---- BPRMF result without injecting pop:")
        print_result_f(ret)
        final["main_branch"] = ret
        print("best_pop_expo in training:", best_pop_expo_normal)
        evaluation_model.set_testing_popularity(np.power(last_stage_popualarity_ori, best_pop_expo_normal))
        ret_BEST_K = evaluation_model.eval(model, None, rec_type='main_with_pop')
        print("|||---BPRMF-A with injecting last stage pop(best gamma):")
        print_result_f(ret_BEST_K)
        final["main_with_pop_last"] = ret_BEST_K
        evaluation_model.set_testing_popularity(np.power(linear_predict_popularity_ori, best_pop_expo_normal))
        ret_BEST_K = evaluation_model.eval(model, None, rec_type='main_with_pop')
        print("|||---BPRMF-A with injecting linear predicted pop (best gamma):")
        print_result_f(ret_BEST_K)
        final["main_with_pop_linear"] = ret_BEST_K
        print("----------------------------")
    elif args.test == 'temp_pop':
        evaluation_model.set_testing_popularity(None)
        ret = evaluation_model.eval(model, None, rec_type='main_branch')
        print("---- result with last pop bias for temp_pop model:")
        print_result_f(ret)
        final["main_branch"] = ret
    else:
        raise NotImplementedError("not implement this test method:" + args.test)
    print("training and testing end!!!!")

    print("|||  ------------------------ best performance for model selected by PD/PDG/BPRMF ------------------- |||")
    print("main best epoch:", config_main['best_epoch'])
    model.restore(save_ckpt_dir + "best_main_ckpt.ckpt")
    print("best result without injecting pop:")
    evaluation_model.set_testing_popularity(None)
    ret = evaluation_model.eval(model, None, rec_type='main_branch')
    print("---- result without injecting pop:")
    print_result_f(ret)
    final["best_main"] = ret
    return model, final
