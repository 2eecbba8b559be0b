"""Host-side tests of the UCI experiment harness (gp_experiment_runner.py): fold arithmetic, train-set
normalisation, the .mat loader, dataset presets, the flag -> solver-setting map and the CSV schema.  The training routine is
replaced by a stub here (the real one needs a GPU: tests/test_model_gpu.py::test_experiment_runner_end_to_end)."""
import json
import os

import numpy as np
import pandas as pd
import pytest
import torch

import gp_experiment_runner as runner
from rpgp.gp import settings


def _frame(n=103, d=4, seed=0):
    rng = np.random.RandomState(seed)
    data = rng.randn(n, d + 1) * np.arange(1, d + 2) + 3.0
    return runner.frame_from_array(data), data


def test_frame_has_reference_schema_and_standardised_target():
    df, data = _frame()
    assert list(df.columns) == ["index", "0", "1", "2", "3", "target"]
    assert abs(df["target"].mean()) < 1e-12 and abs(df["target"].std() - 1.0) < 1e-12
    np.testing.assert_allclose(df["2"].values, data[:, 2])
    data_nan = data.copy()
    data_nan[:, 1] = np.nan
    assert "1" not in runner.frame_from_array(data_nan).columns          # all-NaN columns are dropped
    assert runner.feature_columns(df) == ["0", "1", "2", "3"]


def test_load_dataset_reads_the_reference_mat_layout(tmp_path, monkeypatch):
    from scipy.io import savemat
    _, data = _frame(n=40)
    os.makedirs(tmp_path / "uci" / "toy")
    savemat(str(tmp_path / "uci" / "toy" / "toy.mat"), {"data": data})
    monkeypatch.setattr(runner, "data_base_path", str(tmp_path))
    df = runner.load_dataset("toy")
    assert len(df) == 40 and list(df.columns)[-1] == "target"
    np.testing.assert_allclose(df["0"].values, data[:, 0])


@pytest.mark.parametrize("n,split,expected", [(103, 0.1, [11, 11, 11] + [10] * 7), (100, 0.1, [10] * 10), (105, 0.3, [32, 32, 32]),
                                              (17, 0.5, [9, 8])])
def test_fold_boundaries(n, split, expected):
    df, _ = _frame(n=n)
    starts = runner._determine_folds(split, df)
    assert starts[0] == 0 and list(np.diff(starts)) == expected
    seen = []
    for fold in range(len(starts) - 1):
        train, test = runner._access_fold(df, starts, fold)
        assert len(train) + len(test) == n and set(train["index"]).isdisjoint(test["index"])
        seen += list(test["index"])
    assert seen == list(range(sum(expected)))           # contiguous folds in order; the tail never becomes a test row


def test_normalisation_uses_training_statistics_only():
    df, _ = _frame(n=60)
    df["3"] = 7.0                                        # constant feature: centred, not divided
    starts = runner._determine_folds(0.25, df)
    train, test = runner._access_fold(df, starts, 1)
    ntrain, ntest = runner._normalize_by_train(train, test)
    cols = ["0", "1", "2", "target"]
    np.testing.assert_allclose(ntrain[cols].mean().values, 0.0, atol=1e-12)
    np.testing.assert_allclose(ntrain[cols].std().values, 1.0, atol=1e-12)
    np.testing.assert_allclose(ntrain["3"].values, 0.0)
    mu, sd = train["0"].mean(), train["0"].std()
    np.testing.assert_allclose(ntest["0"].values, (test["0"].values - mu) / sd)
    np.testing.assert_array_equal(ntrain["index"].values, train["index"].values)     # the index column is left alone


def test_dataset_presets():
    everything = runner.get_datasets()
    assert len(everything) == 36 and everything[0] == "challenger" and "song" in runner.get_big_datasets()
    assert runner.resolve_datasets(["all"]) == everything
    assert runner.resolve_datasets(["small"]) == runner.get_small_datasets()
    assert runner.resolve_datasets(["small-med"]) == everything[:18] and runner.resolve_datasets(["small-med"])[-1] == "wine"
    assert runner.resolve_datasets(["med"]) == everything[18:24] and runner.resolve_datasets(["large"]) == everything[24:]
    assert runner.resolve_datasets(["3"]) == everything[:3]
    assert runner.resolve_datasets(["song"]) == ["song"] and runner.resolve_datasets(["a", "b"]) == ["a", "b"]


def test_flags_map_onto_solver_settings():
    args = runner.build_parser().parse_args(["-m", "x.json", "-d", "song", "-o", "out.csv"])
    assert (args.cg_tol, args.eval_cg_tol, args.split, args.max_cg_iterations, args.cv) == (0.05, 0.01, 0.1, 10_000, True)
    args = runner.build_parser().parse_args(["-m", "x.json", "-d", "song", "-o", "o.csv", "--cg_tol", "0.002", "--eval_cg_tol", "0.0005",
                                             "--fast_pred", "--use_chol", "--max_cg_iterations", "77", "--skip_log_det_forward"])
    before = settings.cg_tolerance.value()
    with runner.solver_settings(args):
        assert settings.cg_tolerance.value() == 0.002 and settings.eval_cg_tolerance.value() == 0.0005
        assert settings.fast_pred_var.on() and settings.max_cg_iterations.value() == 77 and settings.skip_logdet_forward.on()
        assert not settings.fast_computations.solves.on() and not settings.fast_computations.log_prob.on()
    assert settings.cg_tolerance.value() == before and not settings.fast_pred_var.on()
    routine, options = runner.routine_and_options({"kind": "rp_poly", "model_kwargs": {}, "train_kwargs": {}}, args)
    assert options["devices"] == ["cuda:0"] and options["evaluate_on_train"] and not options["skip_posterior_variances"]
    with pytest.raises(NotImplementedError):
        runner.routine_and_options({"kind": "ppr_gp"}, args)
    bad = runner.build_parser().parse_args(["-m", "x", "-d", "s", "-o", "o", "--record_pred_unc", "--skip_posterior_variances"])
    with pytest.raises(ValueError):
        runner.routine_and_options({"kind": "rp_poly"}, bad)


def _stub_routine(calls, fail_first=0):
    def routine(trainX, trainY, testX, testY, **options):
        calls.append((trainX.shape, testX.shape, json.loads(json.dumps(options))))
        if len(calls) <= fail_first:
            raise RuntimeError("transient failure")
        assert trainX.dtype == torch.float32 and trainX.is_contiguous()
        return {"prior_train_nmll": 1.5, "trained_epochs": 3}, torch.zeros_like(testY), None
    return routine


def test_run_experiment_rows_and_retry():
    df, _ = _frame(n=50)
    calls = []
    res = runner.run_experiment(_stub_routine(calls), {"kind": "x"}, df, split=0.2, cv=True, repeats=2, print_to_console=False,
                                addl_metrics={"mae": lambda p, y: float((p - y).abs().mean())})
    assert len(res) == 10 and sorted(set(res["fold"])) == [0, 1, 2, 3, 4] and set(res["repeat"]) == {0, 1}
    for col in ("n", "d", "mse", "rmse", "train_time", "prior_train_nmll", "trained_epochs", "mae"):
        assert col in res.columns
    assert (res["n"] == 50).all() and (res["d"] == 4).all() and calls[0][0] == (40, 4) and calls[0][1] == (10, 4)
    np.testing.assert_allclose(res["rmse"].values ** 2, res["mse"].values)
    # single fold, first two attempts fail: two error rows, then the result
    calls = []
    res = runner.run_experiment(_stub_routine(calls, fail_first=2), {}, df, split=0.2, cv=False, chosen_fold=3, print_to_console=False)
    assert len(res) == 3 and res["error"].notna().sum() == 2 and np.isnan(res["rmse"].values[:2]).all()
    assert (res["fold"] == 3).all() and np.isfinite(res["rmse"].values[2])
    # a routine that always fails gives up after error_repeats attempts
    calls = []
    res = runner.run_experiment(_stub_routine(calls, fail_first=99), {}, df, split=0.5, cv=False, error_repeats=3, print_to_console=False)
    assert len(calls) == 3 and len(res) == 3


def test_main_writes_the_reference_csv_schema(tmp_path, monkeypatch):
    spec = {"kind": "rp_poly", "model_kwargs": {"J": 4, "k": 1}, "train_kwargs": {"optimizer": "adam"}}
    spec_path, out = tmp_path / "spec.json", tmp_path / "out.csv"
    spec_path.write_text(json.dumps(spec))
    calls = []
    monkeypatch.setattr(runner.training_routines, "train_exact_gp", _stub_routine(calls))
    df, _ = _frame(n=40)
    table = runner.main(["-m", str(spec_path), "-d", "toy", "-o", str(out), "--no_cv", "--fold", "1", "--ablation", "--J", "2", "5",
                         "--cg_tol", "0.01"], datasets_override={"toy": df})
    assert len(table) == 2 and list(table["J"]) == [2, 5] and [c[2]["model_kwargs"]["J"] for c in calls] == [2, 5]
    saved = pd.read_csv(out)
    for col in ("fold", "repeat", "n", "d", "mse", "rmse", "train_time", "dataset", "options", "J") + runner.RUN_COLUMNS:
        assert col in saved.columns, col
    assert (saved["dataset"] == "toy").all() and (saved["cg_tol"] == 0.01).all() and (saved["fold"] == 1).all()
    assert json.loads(saved["options"][0])["kind"] == "rp_poly"


def test_reference_known_answers_for_the_experiment_helpers():
    """the reference's own TestExperimentHelpers (test.py:494-529), same inputs, same expected values"""
    df = pd.DataFrame({"index": [0, 1, 2], "0": [1, 2, 2], "target": [0, 0, 1]})
    new_train, new_test = runner._normalize_by_train(df.iloc[:2, :], df.iloc[2:, :])
    mean, std = 0.5, np.std([-0.5, 0.5], ddof=1)
    assert new_train["0"].iloc[0] == (0 - mean) / std and new_test["0"].iloc[0] == (0 + mean) / std
    assert new_test["target"].iloc[0] == 1 and new_train["target"].iloc[0] == 0        # zero-variance target: centred only
    df4 = pd.DataFrame({"index": [0, 1, 2, 3], "0": [1, 2, 2, 4], "target": [0, 0, 1, 1]})
    starts = runner._determine_folds(1 / 3, df4)
    assert len(starts) == 4 and starts == [0, 2, 3, 4] and all(type(v) is int for v in starts)
    train, test = runner._access_fold(df4, [0, 2, 3, 4], 0)
    assert test["index"].values.tolist() == [0, 1] and train["index"].values.tolist() == [2, 3]
    train, test = runner._access_fold(df4, [0, 2, 3, 4], 1)
    assert test["index"].values.tolist() == [2] and train["index"].values.tolist() == [0, 1, 3]
