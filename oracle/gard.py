"""Oracle: GARD analog models for ONE cell (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows skdownscale/pointwise_models/gard.py:58-87 (AnalogBase.fit),
:273-364 (PureAnalog.predict) and :152-224 (AnalogRegression.predict /
_predict_one_step, with and without ``thresh``).  The neighbour search of
sklearn.neighbors.KDTree (scikit-learn, pinned 1.7.2 in the reference's
uv.lock:3033-3034; not vendored) is restated as an exact float64 brute force:
squared Euclidean distance accumulated feature by feature, neighbours in
ascending distance, lowest train index first on exact ties (KDTree's own tie
order is implementation-defined, so parity inputs are tie-free).
"""

from __future__ import annotations

import numpy as np


def knn_bruteforce(X_train: np.ndarray, X_query: np.ndarray, k: int):
    """``KDTree(X_train).query(X_query, k)`` (gard.py:82,194,299).

    Returns (dist float64 [Tq,k], inds int64 [Tq,k]).  Distances are
    ``sqrt(sum_j (q_j - a_j)**2)`` with the sum taken j = 0..p-1 in order, in
    float64 (KDTree copies its data to float64)."""
    A = np.asarray(X_train, dtype=np.float64)
    Q = np.asarray(X_query, dtype=np.float64)
    if A.ndim == 1:
        A = A[:, None]
    if Q.ndim == 1:
        Q = Q[:, None]
    Tq = Q.shape[0]
    dist = np.empty((Tq, k), dtype=np.float64)
    inds = np.empty((Tq, k), dtype=np.int64)
    for i in range(Tq):
        d2 = np.zeros(A.shape[0], dtype=np.float64)
        for j in range(A.shape[1]):
            diff = Q[i, j] - A[:, j]
            d2 = d2 + diff * diff
        order = np.argsort(d2, kind='stable')[:k]
        inds[i] = order
        dist[i] = np.sqrt(d2[order])
    return dist, inds


def pure_analog_predict(X_train, y_train, X_query, n_analogs: int = 200, kind: str = 'best_analog',
                        thresh=None, rand_inds=None, return_inds: bool = False):
    """PureAnalog.fit + predict (gard.py:58-87, 273-364).

    Returns float64 [Tq, 3] in ``output_names`` order (pred, exceedance_prob,
    prediction_error); the individual columns are first computed in the dtypes
    the reference produces (pred/err in y's dtype for best/mean, float64 for
    weight) and then widened exactly.  ``rand_inds`` replaces the global-RNG draw
    ``np.random.randint(0, k, size=Tq)`` of gard.py:315 for 'sample_analogs'.
    """
    y_ = np.asarray(y_train).reshape(-1)
    Tq = len(np.asarray(X_query))
    k_ = min(n_analogs, len(y_))                                    # gard.py:75-79
    if kind == 'best_analog' or n_analogs == 1:                     # gard.py:290-296
        k, kind = 1, 'best_analog'
    else:
        k = k_
    dist, inds = knn_bruteforce(X_train, X_query, k)                # gard.py:299
    analogs = np.take(y_, inds, axis=0)                             # gard.py:301
    if thresh is not None:                                          # gard.py:303-308
        analog_mask = analogs > thresh
        masked = np.where(analog_mask, analogs, np.nan)
    if kind == 'best_analog':
        predicted = analogs[:, 0]
    elif kind == 'sample_analogs':
        if rand_inds is None:
            raise ValueError('sample_analogs needs the host-drawn rand_inds')
        predicted = analogs[np.arange(Tq), np.asarray(rand_inds)].astype(np.float64)   # gard.py:19-24
    elif kind == 'weight_analogs':
        weights = 1.0 / np.where(dist == 0, 1e-20, dist)            # gard.py:322-323
        src = masked if thresh is not None else analogs
        predicted = np.average(src, weights=weights, axis=1)        # gard.py:324-327
    elif kind == 'mean_analogs':
        predicted = (masked if thresh is not None else analogs).mean(axis=1)   # gard.py:329-333
    else:
        raise ValueError(f'got unexpected kind {kind}')
    if thresh is not None:                                          # gard.py:338-343
        predicted = np.nan_to_num(predicted, nan=0.0)
        prediction_error = masked.std(axis=1)
        exceedance_prob = np.where(analog_mask, 1, 0).mean(axis=1)
    else:                                                           # gard.py:344-346
        prediction_error = analogs.std(axis=1)
        exceedance_prob = np.ones(Tq, dtype=np.float64)
    out = np.stack([np.asarray(predicted, dtype=np.float64),
                    np.asarray(exceedance_prob, dtype=np.float64),
                    np.asarray(prediction_error, dtype=np.float64)], axis=1)
    return (out, inds, dist) if return_inds else out


def logistic_fit_exact(x: np.ndarray, b: np.ndarray, C: float = 1.0):
    """The optimum sklearn's ``LogisticRegression(C=C)`` (l2 penalty, intercept not penalised,
    lbfgs) iterates towards (gard.py:172,209):  minimise
    ``sum_i log(1 + exp(-s_i (w.x_i + c))) + ||w||^2 / (2 C)``, ``s_i = +-1``.
    scikit-learn (pinned 1.7.2, uv.lock:3033-3034; not vendored) stops lbfgs at a projected
    gradient of ``tol = 1e-4`` on the per-sample-averaged objective, so its answer is this optimum
    only to ~1e-4..1e-3; here damped Newton runs to machine precision.  Returns (w, c)."""
    x = np.asarray(x, dtype=np.float64)
    t = np.asarray(b, dtype=np.float64)
    n, p = x.shape
    Z = np.hstack([x, np.ones((n, 1))])
    th = np.zeros(p + 1)
    reg = np.r_[np.full(p, 1.0 / C), 0.0]

    def objective(v):
        z = Z @ v
        return np.sum(np.logaddexp(0.0, z) - t * z) + 0.5 * np.sum(reg * v * v)

    f = objective(th)
    for _ in range(200):
        z = Z @ th
        mu = 1.0 / (1.0 + np.exp(-z))
        g = Z.T @ (mu - t) + reg * th
        H = (Z * (mu * (1.0 - mu))[:, None]).T @ Z + np.diag(reg)
        d = np.linalg.solve(H + 1e-300 * np.eye(p + 1), -g)
        step = 1.0
        while True:
            cand = th + step * d
            fc = objective(cand)
            if fc <= f + 1e-4 * step * (g @ d) or step < 1e-10:
                break
            step *= 0.5
        th, f_old, f = cand, f, fc
        if np.max(np.abs(step * d)) < 1e-14 * max(1.0, np.max(np.abs(th))):
            break
    return th[:p], th[p]


def analog_regression_predict(X_train, y_train, X_query, n_analogs: int = 200, thresh=None,
                              return_inds: bool = False, logistic_C: float = 1.0):
    """AnalogRegression.fit + predict (gard.py:152-224).

    Per timestep: k nearest analogs → [thresh] logistic regression of ``y > thresh`` on the
    analog predictors, ``exceedance_prob = predict_proba[0, 0]`` = P(class 0) (gard.py:201-212; 1.0
    when every analog exceeds; ValueError when none does, like sklearn) → ordinary least squares
    with intercept on the (exceeding) float64 analog predictors (sklearn LinearRegression = centred
    ``lstsq``, minimum-norm when rank deficient) → prediction at the query point,
    prediction_error = in-sample RMSE."""
    A = np.asarray(X_train, dtype=np.float64)
    if A.ndim == 1:
        A = A[:, None]
    Q = np.asarray(X_query)
    if Q.ndim == 1:
        Q = Q[:, None]
    y_raw = np.asarray(y_train).reshape(-1)
    k_ = min(n_analogs, len(y_raw))
    _, inds = knn_bruteforce(A, Q, k_)                              # gard.py:194
    out = np.empty((Q.shape[0], 3), dtype=np.float64)
    for i in range(Q.shape[0]):
        x = A[inds[i]]                                              # gard.py:197
        yk = y_raw[inds[i]]                                         # gard.py:198
        prob = 1.0
        if thresh is not None:
            exceed = yk > thresh                                    # gard.py:201-202 (in y's dtype)
            if not exceed.all():                                    # gard.py:208-212
                if not exceed.any():
                    raise ValueError('This solver needs samples of at least 2 classes in the data, but the data '
                                     'contains only one class: 0')
                w, c0 = logistic_fit_exact(x, exceed, logistic_C)
                z = float(Q[i].astype(np.float64) @ w + c0)
                prob = 1.0 - 1.0 / (1.0 + np.exp(-z))               # predict_proba[0, 0] = P(class 0)
            x, yk = x[exceed], yk[exceed]
        y = yk.astype(np.float64)                                   # sklearn casts the target
        xo = x.mean(axis=0)
        yo = y.mean()
        coef, *_ = np.linalg.lstsq(x - xo, y - yo, rcond=None)      # gard.py:215
        intercept = yo - xo @ coef
        y_hat = x @ coef + intercept                                # gard.py:218
        err = np.sqrt(np.mean((y - y_hat) ** 2))                    # gard.py:219
        pred = Q[i].astype(np.float64) @ coef + intercept           # gard.py:221
        out[i] = (pred, prob, err)                                  # gard.py:224
    return (out, inds) if return_inds else out


def pure_regression_fit_predict(X_train, y_train, X_query, thresh=None, logistic_C: float = 1.0):
    """PureRegression.fit + predict for ONE cell (gard.py:367-504) — SURVEY §8(f) row 4.

    One least-squares fit per cell on the rows whose target exceeds ``thresh`` (all rows without a
    threshold), in the dtype of the inputs like sklearn's LinearRegression (float32 inputs are fitted
    in float32 — LAPACK details differ between drivers at the 1e-6 level, so float32 parity is to the
    tolerance, float64 parity to 1e-9); ``prediction_error`` = the in-sample RMSE, the same for every
    step; ``exceedance_prob`` = P(class 1) of a logistic regression of ``y > thresh`` on ALL rows
    (gard.py:417, 467 — note AnalogRegression reports P(class 0)), 1.0 without a threshold or when
    every row exceeds (gard.py:418-430: the one-class case mutates ``thresh`` to None; with no row
    above the threshold the linear fit then fails on an empty set and the reference raises).
    Returns float64 [Tq, 3] in ``output_names`` order."""
    X = np.asarray(X_train)
    if X.ndim == 1:
        X = X[:, None]
    y = np.asarray(y_train).reshape(-1)
    Q = np.asarray(X_query)
    if Q.ndim == 1:
        Q = Q[:, None]
    prob = np.ones(len(Q), dtype=np.float64)
    exceed = np.ones(len(y), dtype=bool)
    if thresh is not None:
        exceed = y > thresh
        if exceed.all() or not exceed.any():
            if not exceed.any():
                raise ValueError('Found array with 0 sample(s) (shape=(0, %d)) while a minimum of 1 is required '
                                 'by LinearRegression.' % X.shape[1])
        else:
            w, c0 = logistic_fit_exact(X.astype(np.float64), exceed, logistic_C)
            z = Q.astype(np.float64) @ w + c0
            prob = 1.0 / (1.0 + np.exp(-z))                          # predict_proba[:, 1]
    xs, ys = X[exceed], y[exceed]
    xo, yo = xs.mean(axis=0), ys.mean()
    coef, *_ = np.linalg.lstsq(xs - xo, ys - yo, rcond=None)
    intercept = yo - xo @ coef
    resid = ys - (xs @ coef + intercept)
    err = float(np.sqrt(np.mean(resid.astype(np.float64) ** 2)))
    pred = Q @ coef + intercept
    return np.stack([pred.astype(np.float64), prob, np.full(len(Q), err)], axis=1)
