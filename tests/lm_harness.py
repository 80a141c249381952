"""TEST INFRASTRUCTURE: the Levenberg-Marquardt state machine of CorrelationClass::Newton_Raphson
(correlation_class.cpp:349-640) in plain Python, with the evaluation (A, b, chi of one pass over the level's pixels)
and the damped solve as pluggable callables. Lets a test mix the device's evaluation with the oracle's solver (and the
other way round) to attribute a difference in the final chi to its source, and prints the trace of decisions."""
import numpy as np

F = np.float32


def translate(p, src, dst, n_params):
    """pyramid_class.cpp:260-287: only u, v scale between levels (second-order terms of the 12-parameter extension
    scale the other way)."""
    q = np.array(p, F)
    mag = F(1.0) / F(1 << (dst - src)) if dst > src else F(1 << (src - dst))
    q[:2] *= mag
    if n_params == 12:
        q[6:] *= F(1.0) / mag
    return q


def newton_raphson(evaluate, solve, points_per_level, guess, pyramid, precision=1e-3, max_iters=50, trace=None):
    """evaluate(level, p) -> (A_upper (n x n), b (n), chi_sum, out_of_image); solve(A_upper, b, lam, scaling) -> dp.
    Returns dict(params, chi, iterations, evaluations[level], error_code)."""
    start, step, stop = pyramid
    n = len(guess)
    precision = F(precision)
    min_lambda, max_lambda = F(1e-9), F(1e9)
    mp = translate(guess, 0, stop, n)
    level_old = 0
    evals = {}
    error_code, reached = 0, 0
    last_good_chi = F(np.finfo(np.float32).max)
    level = stop
    while level >= start:
        if level != stop:
            mp = translate(mp, level_old, level, n)
        error_code = 0
        lam = F(1e-4)
        last_good_chi = F(np.finfo(np.float32).max)
        last_good = mp.copy()
        scaling = F(1.0) / F(points_per_level[level])
        evals[level] = 0

        def ev(p, what):
            A, b, chi_sum, oob = evaluate(level, p)
            evals[level] += 1
            chi = F(chi_sum) * scaling
            if trace is not None:
                trace.append((level, what, float(chi), float(lam), np.array(p, F).copy()))
            return A, b, chi, oob

        # INIT (:410-439)
        A, b, chi, oob = ev(mp, "init")
        if oob:
            error_code = 2
            level_old = level
            break
        last_good_chi = chi
        saved = (mp + solve(A, b, lam, scaling)).astype(F)
        mp = saved.copy()
        use_saved, iteration = True, 1
        while True:  # :441-585
            if iteration > max_iters or lam >= max_lambda:
                error_code = 3
                break
            reached = iteration
            if use_saved:
                tentative = saved.copy()
            else:
                A, b, chi, oob = ev(last_good, "redo")
                if oob:
                    error_code, mp = 2, last_good.copy()
                    break
                tentative = (last_good + solve(A, b, lam, scaling)).astype(F)
                mp = tentative.copy()
            A, b, chi, oob = ev(tentative, "tent")
            if oob:
                error_code, mp = 2, tentative.copy()
                break
            saved = (tentative + solve(A, b, max(lam * F(0.4), min_lambda), scaling)).astype(F)
            mp = saved.copy()
            delta = abs((last_good_chi - chi) / (max(last_good_chi, chi) + precision))
            if chi <= last_good_chi:
                last_good_chi, lam, last_good, use_saved = chi, max(lam * F(0.4), min_lambda), tentative.copy(), True
                if trace is not None:
                    trace.append((level, "accept", float(chi), float(lam), None))
            else:
                lam, use_saved = min(lam * F(10.0), max_lambda), False
                if trace is not None:
                    trace.append((level, "reject", float(chi), float(lam), None))
            if delta < precision:
                break
            iteration += 1
        level_old = level
        level -= step
    params = translate(mp, level_old, 0, n)
    return dict(params=params, chi=F(last_good_chi), iterations=reached, evaluations=evals, error_code=error_code)
