"""The optimisation loop that drives the K.V path: closure -> model(x) -> -MLL -> backward -> optimizer step.

API mirror of the reference's fitting/optimizing.py: `train_to_convergence` (:14-108, same signature and defaults,
moving-average convergence test, best-state checkpointing) and `mean_squared_error` (:111-113).  `learn_projections`
(PPR backfitting, :116-160) belongs to a baseline outside the hot path.
"""
import copy
import gc
from typing import Optional, Type

import numpy as np
import torch


def _batches(xs, ys, batch_size):
    """full batch by default (what every spec uses); shuffled mini-batches when batch_size is given"""
    n = xs.shape[0]
    if batch_size is None or batch_size >= n:
        yield xs, ys
        return
    perm = torch.randperm(n, device=xs.device)
    for i in range(0, n, batch_size):
        idx = perm[i:i + batch_size]
        yield xs[idx], ys[idx]


def train_to_convergence(model, xs, ys, optimizer: Optional[Type] = None, lr=0.1, objective=None, max_iter=100, verbose=0,
                         patience=20, conv_tol=1e-4, check_conv=True, smooth=True, isloss=False, batch_size=None,
                         checkpoint=False, print_freq=1):
    """Optimise `objective` (maximised unless `isloss`) over the model's parameters.

    Stops after max_iter epochs, or -- when check_conv -- once the (moving-average, if `smooth`) loss has improved by
    less than conv_tol over the last `patience` epochs.  With `checkpoint` the best-loss state is restored on exit.
    Returns the number of epochs run (the epoch index at convergence, max_iter otherwise).
    """
    if optimizer is None:
        optimizer = torch.optim.LBFGS
    verbose = int(verbose)
    model.train()
    optimizer_ = optimizer(model.parameters(), lr=lr)
    gc.collect()

    best_state = None
    best_loss = np.inf
    losses = np.zeros((max_iter,))
    ma = np.zeros((max_iter,))

    def finish(epochs):
        if checkpoint and best_state is not None:
            model.load_state_dict(best_state)
        return epochs

    for i in range(max_iter):
        total_loss = 0
        for j, (x_batch, y_batch) in enumerate(_batches(xs, ys, batch_size)):
            def closure():  # LBFGS re-evaluates; Adam/SGD call it once
                optimizer_.zero_grad()
                value = objective(model(x_batch), y_batch)
                loss = value if isloss else -value
                loss.backward()
                return loss
            loss = optimizer_.step(closure).item()
            if verbose > 1:
                print("epoch {}, iter {}, loss {}".format(i, j, loss))
            total_loss = total_loss + loss
        losses[i] = total_loss
        # moving average over the last `patience` epochs; an empty slice (early epochs) gives NaN, exactly like the
        # reference (:81), so the smooth criterion cannot fire before epoch 2*patience-1
        with np.errstate(all="ignore"):
            window = losses[max(i - patience + 1, 0):i + 1] if i - patience + 1 >= 0 else losses[0:0]
            ma[i] = window.mean() if window.size else np.nan
        if verbose >= 1 and i % print_freq == 0:
            print("epoch {}, loss {}, noise {}".format(i, total_loss, model.likelihood.noise.item()))
        if checkpoint and total_loss < best_loss:
            best_loss = total_loss
            best_state = copy.deepcopy(model.state_dict())
        if check_conv and i >= patience:
            if smooth and ma[i - patience] - ma[i] < conv_tol:
                if verbose > 0:
                    print("Reached convergence at {}, MA {} - {} < {}".format(total_loss, ma[i - patience], ma[i], conv_tol))
                return finish(i)
            if not smooth and losses[i - patience] - losses[i] < conv_tol:
                if verbose > 0:
                    print("Reached convergence at {}, {} - {} < {}".format(total_loss, losses[i - patience], total_loss, conv_tol))
                return finish(i)
    return finish(max_iter)


def mean_squared_error(y_pred, y_true):
    return ((y_pred - y_true) ** 2).mean().item()
