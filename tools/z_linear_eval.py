#!/usr/bin/env python
"""Downstream consumer of the compressed representations: linear evaluation of
decompressed Z, the "lossless prediction" half of the reference's pipeline -- README.md:66-74
(`decompress_dataset` -> `LinearSVC(C=7e-3).fit`) and utils/Z_linear_eval.py:62-93 (randomised search over
C / class_weight, accuracy + balanced accuracy).  Hydra / PretrainedAnalyser orchestration is out of scope;
this takes the container files `compress_dataset` wrote.

    python tools/z_linear_eval.py train.bin train_y.npy test.bin test_y.npy [--beta 0.05] [--search 0]
"""
import argparse
import json
import os
import sys
import time
import warnings


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def linear_eval(Z_train, Y_train, Z_test, Y_test, n_search=0, seed=123, C=7e-3):
    """README recipe (fixed C) or utils/Z_linear_eval.py's RandomizedSearchCV when n_search > 0."""
    from sklearn.metrics import accuracy_score, balanced_accuracy_score
    from sklearn.svm import LinearSVC

    t0 = time.time()
    if n_search > 0:
        from scipy.stats import loguniform
        from sklearn.model_selection import RandomizedSearchCV

        clf = RandomizedSearchCV(LinearSVC(random_state=seed, dual=False),
                                 {"C": loguniform(1e-3, 1e-0), "class_weight": ["balanced", None]},
                                 scoring="accuracy", n_jobs=8, cv=5, n_iter=n_search, random_state=seed)
    else:
        clf = LinearSVC(C=C, random_state=seed, dual=False)
    clf.fit(Z_train, Y_train)
    Y_pred = clf.predict(Z_test)
    acc, bacc = accuracy_score(Y_test, Y_pred), balanced_accuracy_score(Y_test, Y_pred)
    return {"test/pred/acc": acc, "test/pred/balanced_acc": bacc, "test/pred/err": 1 - acc,
            "test/pred/balanced_err": 1 - bacc, "fit_seconds": time.time() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("train_file"), ap.add_argument("train_labels"), ap.add_argument("test_file"), ap.add_argument("test_labels")
    ap.add_argument("--beta", type=float, default=0.05, choices=[0.01, 0.05, 0.1])
    ap.add_argument("--search", type=int, default=0, help="RandomizedSearchCV iterations (0 = README's fixed C)")
    a = ap.parse_args()
    warnings.filterwarnings("ignore", message="no CLIP weights given")
    import hubconf

    comp, _ = {0.01: hubconf.clip_compressor_b001, 0.05: hubconf.clip_compressor_b005, 0.1: hubconf.clip_compressor_b01}[a.beta]()
    Z_train, Y_train = comp.decompress_dataset(a.train_file, a.train_labels)
    Z_test, Y_test = comp.decompress_dataset(a.test_file, a.test_labels)
    print(json.dumps(linear_eval(Z_train, Y_train, Z_test, Y_test, a.search)))


if __name__ == "__main__":
    main()
