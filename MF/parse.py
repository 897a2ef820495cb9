"""Command-line flags of MF/train_new_api.py -- the reference's MF/parse.py:3-117 flag set, kept name for name
and default for default (flags the reference marks "not used" still parse and are ignored)."""
import argparse

# (flag, type, default, help)
_FLAGS = [
    ("data_path", str, "./data/", "Input data path (popularity table; interaction files come from ./data/<dataset>/)."),
    ("dataset", str, "kwai", "Dataset directory name."),
    ("source", str, "normal", "not used"),
    ("train", str, "normal", "normal (BPRMF) | s_condition (PD/PDA) | condition (PDG) | temp_pop (BPR(t)-pop)"),
    ("test", str, "normal", "normal | s_condition | condition | temp_pop"),
    ("valid_set", str, "test", "test | valid"),
    ("save_dir", str, "/data/zyang/save_model/", "checkpoint root"),
    ("alpha", float, 1e-3, "not used (appears in the checkpoint path)"),
    ("beta", float, 1e-3, "not used"),
    ("pc_alpha", float, 0.1, "not used"),
    ("pc_beta", float, 0.1, "not used"),
    ("exp_init_values", float, 0.1, "not used"),
    ("pop_exp", float, 0.1, "popularity power coefficient (gamma)"),
    ("early_stop", int, 1, "1: stop when both criteria stalled for 100 // log_interval evaluations"),
    ("need_save", int, 1, "not used"),
    ("cores", int, 1, "not used (the sampler runs on the GPU)"),
    ("verbose", int, 1, "print the epoch loss every `verbose` epochs between evaluations"),
    ("epoch", int, 400, "number of epochs"),
    ("load_epoch", int, 400, "not used"),
    ("embed_size", int, 64, "embedding size d"),
    ("batch_size", int, 1024, "triples per step"),
    ("Ks", str, "[20]", "cut-offs of the metrics, e.g. \"[20,50]\" (top-50 is kept)"),
    ("epochs", str, "[]", "not used"),
    ("regs", float, 1e-5, "L2 coefficient"),
    ("fregs", float, 1e-5, "not used"),
    ("c", float, 10.0, "not used"),
    ("train_c", str, "val", "not used"),
    ("lr", float, 1e-3, "Adam learning rate"),
    ("wd", float, 1e-5, "not used (overwritten by regs for the checkpoint path)"),
    ("model", str, "mf", "mf"),
    ("skew", int, 0, "not used"),
    ("model_type", str, "o", "not used"),
    ("devide_ratio", float, 0.8, "not used"),
    ("save_flag", int, 1, "1: also checkpoint every 50 epochs"),
    ("pop_used", int, -2, "not used"),
    ("cuda", str, "1", "CUDA_VISIBLE_DEVICES"),
    ("pretrain", int, 0, "not used"),
    ("check_c", int, 1, "not used"),
    ("log_interval", int, 10, "evaluate every `log_interval` epochs"),
    ("pop_wd", float, 0.0, "not used"),
    ("base", float, -1.0, "not used"),
    ("cf_pen", float, 1.0, "not used"),
    ("saveID", str, "", "checkpoint path suffix"),
    ("user_min", int, 1, "not used"),
    ("user_max", int, 1000, "not used"),
    ("data_type", str, "ori", "ori"),
    ("imb_type", str, "exp", "not used"),
    ("top_ratio", float, 0.1, "not used"),
    ("lam", float, 1.0, "not used"),
    ("check_epoch", str, "all", "not used"),
    ("start", float, -1.0, "not used"),
    ("end", float, 1.0, "not used"),
    ("step", int, 20, "not used"),
    ("out", int, 0, "not used"),
]


def build_parser():
    parser = argparse.ArgumentParser(description="Run pop_bias.")
    for name, typ, default, text in _FLAGS:
        if typ is str:
            parser.add_argument("--" + name, nargs="?", default=default, help=text)
        else:
            parser.add_argument("--" + name, type=typ, default=default, help=text)
    return parser


def parse_args(argv=None):
    return build_parser().parse_args(argv)
