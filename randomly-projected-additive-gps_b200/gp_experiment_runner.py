"""UCI experiment harness around the fused K.V path.

Mirror of the reference's `gp_experiment_runner.py` for the exact-GP family: the same public functions (`load_dataset`,
`run_experiment`, the dataset lists, the fold helpers), the same command-line flags and the same CSV columns
(reference gp_experiment_runner.py:17-102 data handling, :105-219 run_experiment, :222-381 CLI), so that the authors' protocol
runs unchanged once the `.mat` files are placed under `config.data_base_path`.  What differs is behind the boundary: the solver
settings map onto `rpgp.gp.settings` (there is one backend: the sm_100a kernels; `--device` must name CUDA devices) and only
`kind`s that lower to the additive-RBF operator are accepted.

Layout of a dataset on disk, as in the reference (:21): `<data_base_path>/uci/<name>/<name>.mat` with one array `data`, last
column = target.
"""
import argparse
import contextlib
import datetime
import json
import os
import time
import traceback

import numpy as np
import pandas as pd
import torch

import training_routines
from config import data_base_path
from fitting.optimizing import mean_squared_error
from rpgp.gp import settings as gp_settings

SMALL = ["challenger", "fertility", "concreteslump", "autos", "servo", "breastcancer", "machine", "yacht", "autompg", "housing",
         "forest", "stock", "pendulum", "energy"]
MEDIUM = ["concrete", "solar", "airfoil", "wine", "gas", "skillcraft", "sml", "parkinsons", "pumadyn32nm"]
BIG = ["pol", "elevators", "bike", "kin40k", "protein", "tamielectric", "keggdirected", "slice", "keggundirected", "3droad", "song",
       "buzz", "houseelectric"]
DEFAULT_ABLATION = [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377]


def get_small_datasets():
    return list(SMALL)


def get_medium_datasets():
    return list(MEDIUM)


def get_big_datasets():
    return list(BIG)


def get_datasets():
    return SMALL + MEDIUM + BIG


def frame_from_array(data):
    """(n, d+1) array, target last -> the frame the rest of the harness works on: string feature names '0'..'d-1', an `index`
    column (from reset_index) and a standardised `target`; all-NaN columns dropped (reference :22-33)."""
    data = np.asarray(data)
    d = data.shape[1] - 1
    df = pd.DataFrame(data, columns=[str(i) for i in range(d)] + ["target"]).reset_index()
    t = df["target"] - df["target"].mean()
    df["target"] = t / t.std()
    return df.dropna(axis=1, how="all")


def load_dataset(name):
    from scipy.io import loadmat
    path = os.path.join(data_base_path, "uci", name, name + ".mat")
    return frame_from_array(loadmat(path)["data"])


def format_timedelta(delta):
    hours, rest = divmod(delta.seconds, 3600)
    minutes, seconds = divmod(rest, 60)
    return "%dd %dh %dm %ds" % (delta.days, hours, minutes, seconds)


def feature_columns(frame):
    return [c for c in frame.columns if c != "target" and str(c).lower() != "index"]


def _determine_folds(split, dataset):
    """Start offsets of round(1/split) contiguous folds of floor(n*split) rows; the first `remainder` folds take one row more
    (reference :65-76; rows beyond the last fold always stay in the training part)."""
    n = len(dataset)
    size, folds = int(np.floor(n * split)), int(round(1.0 / split))
    extra = n - size * folds
    sizes = [size + 1 if i < extra else size for i in range(folds)]
    return [0] + [int(v) for v in np.cumsum(sizes)]


def _access_fold(dataset, fold_starts, fold):
    lo, hi = int(fold_starts[fold]), int(fold_starts[fold + 1])
    test = dataset.iloc[lo:hi]
    train = pd.concat([dataset.iloc[:lo], dataset.iloc[hi:]])
    return train, test


def _normalize_by_train(train, test):
    """centre features and target on the training means, divide by the training standard deviations that are non-zero
    (reference :87-102)"""
    train, test = train.copy(), test.copy()
    cols = feature_columns(train) + ["target"]
    mu = train[cols].mean()
    sd = train[cols].std()
    sd = sd.where(sd > 0, 1.0)
    train[cols] = (train[cols] - mu) / sd
    test[cols] = (test[cols] - mu) / sd
    return train, test


def _tensors(frame, features):
    X = torch.tensor(frame[features].values, dtype=torch.float).contiguous()
    y = torch.tensor(frame["target"].values, dtype=torch.float).contiguous()
    return X, y


def run_experiment(training_routine, training_options, dataset, split, cv, addl_metrics=None, repeats=1, error_repeats=10,
                   normalize_using_train=True, chosen_fold=0, print_to_console=True):
    """Train / evaluate `training_routine(trainX, trainY, testX, testY, **training_options)` on the folds of `dataset` (a name or
    an already shuffled frame) and return one result row per (fold, repeat): fold, repeat, n, d, mse, rmse, train_time, every
    entry of the routine's metric dict, every additional metric.  A fold that raises is retried up to `error_repeats` times and
    each failure is logged as a row with the traceback (reference :105-219)."""
    addl_metrics = addl_metrics or {}
    if isinstance(dataset, str):
        dataset = load_dataset(dataset)
    features = feature_columns(dataset)
    starts = _determine_folds(split, dataset)
    n_folds = len(starts) - 1
    rows, t0 = [], time.time()
    for fold in range(n_folds):
        if not cv and fold != chosen_fold:
            continue
        train, test = _access_fold(dataset, starts, fold)
        if normalize_using_train:
            train, test = _normalize_by_train(train, test)
        failures, done = 0, False
        while not done and failures < error_repeats:
            try:
                trainX, trainY = _tensors(train, features)
                testX, testY = _tensors(test, features)
                for repeat in range(repeats):
                    row = {"fold": fold, "repeat": repeat, "n": len(dataset), "d": len(features)}
                    tic = time.perf_counter()
                    out = training_routine(trainX, trainY, testX, testY, **training_options)
                    metrics, ypred = out[0], out[1]
                    row["mse"] = mean_squared_error(ypred, testY)
                    row["rmse"] = float(np.sqrt(row["mse"]))
                    row["train_time"] = time.perf_counter() - tic
                    row.update(metrics)
                    for name, fn in addl_metrics.items():
                        row[name] = fn(ypred, testY)
                    rows.append(row)
                    done = True
                    if print_to_console:
                        finished = fold * repeats + repeat + 1
                        eta = datetime.timedelta(seconds=(time.time() - t0) / finished * (n_folds * repeats - finished))
                        print("%s, fold=%d, rep=%d, eta=%s \n%s" % (datetime.datetime.now(), fold, repeat, format_timedelta(eta), row))
            except Exception:
                failures += 1
                rows.append({"error": traceback.format_exc(), "fold": fold, "n": len(dataset), "d": len(features) - 2,
                             "mse": np.nan, "rmse": np.nan})
                traceback.print_exc()
                print("errors: ", failures)
    results = pd.DataFrame(rows)
    if print_to_console and len(results):
        print("Mean RMSE = {}".format(results["rmse"].mean()))
    return results


# ---- command line --------------------------------------------------------------------------------------------------------------
def build_parser():
    p = argparse.ArgumentParser(description="Run a GP model specification over UCI regression datasets (exact GPs on the fused "
                                            "sm_100a K.V path).")
    p.add_argument("-m", "--model_spec", type=str, required=True, help="path to model specification json file")
    p.add_argument("-d", "--datasets", type=str, nargs="+", required=True, help="UCI dataset name(s) or a predefined set")
    p.add_argument("-o", "--output", type=str, required=True, help="path to output csv file")
    p.add_argument("-s", "--split", type=float, default=0.1, help="fraction of data in test set")
    p.add_argument("-r", "--repeats", type=int, default=1, help="number of times to repeat each fold")
    p.add_argument("--no_cv", action="store_false", dest="cv")
    p.add_argument("--cg_tol", type=float, default=0.05)
    p.add_argument("--eval_cg_tol", type=float, default=0.01)
    p.add_argument("--fast_pred", dest="fast_pred", action="store_true")
    p.add_argument("--use_chol", action="store_true")
    p.add_argument("--no_toeplitz", dest="use_toeplitz", action="store_false")
    p.add_argument("--memory_efficient", dest="memory_efficient", action="store_true")
    p.add_argument("--device", type=str, default="cuda:0", help="CUDA device string(s), comma separated")
    p.add_argument("--skip_posterior_variances", action="store_true")
    p.add_argument("--ablation", action="store_true")
    p.add_argument("--J", type=int, nargs="+", help="Js to use in ablation")
    p.add_argument("--k", type=int, nargs="+", help="ablation over k (coordinates per projection) instead of J")
    p.add_argument("--fold", type=int, default=0)
    p.add_argument("--error_repeats", type=int, default=10)
    p.add_argument("--max_cg_iterations", type=int, default=10_000)
    p.add_argument("--skip_evaluate_on_train", action="store_true")
    p.add_argument("--skip_random_restart", action="store_true")
    p.add_argument("--skip_log_det_forward", action="store_true")
    p.add_argument("--checkpoint_kernel", type=int, default=0, help="accepted for compatibility: the fused operator never forms K")
    p.add_argument("--record_pred_unc", action="store_true")
    p.add_argument("--double", action="store_true", help="double precision (the un-tiled FP64 kernels)")
    return p


def resolve_datasets(names):
    """`all`, `small`, `small-med`, `med`, `large`, an integer N (first N of the full list) or explicit names (reference :267-283)"""
    if len(names) != 1:
        return list(names)
    key, everything = names[0], get_datasets()
    presets = {"all": everything, "small": get_small_datasets(), "small-med": everything[:18], "med": everything[18:24],
               "large": everything[24:]}
    if key in presets:
        return list(presets[key])
    try:
        return everything[:int(key)]
    except ValueError:
        return [key]


def solver_settings(args):
    """the `with gpytorch.settings...` block of the reference (:324-332) as one context manager over rpgp.gp.settings"""
    stack = contextlib.ExitStack()
    fast = not args.use_chol
    for ctx in (gp_settings.cg_tolerance(args.cg_tol), gp_settings.eval_cg_tolerance(args.eval_cg_tol),
                gp_settings.fast_computations(fast, fast, fast), gp_settings.fast_pred_var(args.fast_pred),
                gp_settings.use_toeplitz(args.use_toeplitz), gp_settings.max_cg_iterations(args.max_cg_iterations),
                gp_settings.checkpoint_kernel(args.checkpoint_kernel), gp_settings.skip_logdet_forward(args.skip_log_det_forward),
                gp_settings.memory_efficient(args.memory_efficient)):
        stack.enter_context(ctx)
    return stack


def routine_and_options(options, args):
    """pick the training routine for options['kind'] and merge the flags that travel as keyword arguments (reference :285-319)"""
    options = dict(options)
    kind = options.get("kind")
    if kind in ("ppr_gp", "cgp", "model_average"):
        raise NotImplementedError("kind '%s' is not an exact GP on the fused K.V path (DESIGN.md §9)" % kind)
    options["skip_random_restart"] = args.skip_random_restart
    options["devices"] = args.device.split(",")
    options["skip_posterior_variances"] = args.skip_posterior_variances
    options["evaluate_on_train"] = not args.skip_evaluate_on_train
    options["record_pred_unc"] = args.record_pred_unc
    if args.double:
        options["double"] = True
    if options["record_pred_unc"] and options["skip_posterior_variances"]:
        raise ValueError("Can't record predictive uncertainty while skipping posterior variances.")
    return training_routines.train_exact_gp, options


RUN_COLUMNS = ("cg_tol", "eval_cg_tol", "use_chol", "max_cg_iterations", "use_toeplitz", "fast_pred_var", "checkpoint_kernel",
               "skip_log_det_forward", "memory_efficient")


def main(argv=None, datasets_override=None):
    """`datasets_override`: {name: frame} used instead of the .mat files (tests, synthetic runs)."""
    args = build_parser().parse_args(argv)
    print("Parser arguments", args)
    with open(args.model_spec, "r") as f:
        spec = json.load(f)
    print("Loaded options", spec)
    print("Using device(s) {}".format(args.device.split(",")))
    print("Registered data base path {}".format(data_base_path))
    routine, options = routine_and_options(spec, args)
    if args.ablation:
        key = "k" if args.k is not None else "J"
        values = args.k if args.k is not None else (args.J if args.J is not None else DEFAULT_ABLATION)
    else:
        key, values = None, [-1]
    table = pd.DataFrame()
    for name in resolve_datasets(args.datasets):
        print("Starting dataset {}".format(name))
        data = datasets_override[name] if datasets_override is not None else name
        with solver_settings(args):
            for value in values:
                if key is not None:
                    options["model_kwargs"][key] = value
                results = run_experiment(routine, options, data, split=args.split, cv=args.cv, repeats=args.repeats,
                                         normalize_using_train=True, chosen_fold=args.fold, error_repeats=args.error_repeats)
                if key is not None:
                    results[key] = value
                results["dataset"] = name
                results["options"] = json.dumps(options)
                run_values = (args.cg_tol, args.eval_cg_tol, args.use_chol, args.max_cg_iterations, args.use_toeplitz, args.fast_pred,
                              args.checkpoint_kernel, args.skip_log_det_forward, args.memory_efficient)
                for col, val in zip(RUN_COLUMNS, run_values):
                    results[col] = val
                table = pd.concat([table, results])
                table.to_csv(args.output)
    return table


if __name__ == "__main__":
    main()
