"""Hyper-parameter search with the Hyperopt surface the reference uses::

    from b200ddl.hpo import fmin, hp, tpe, rand, Trials, ParallelTrials, STATUS_OK, STATUS_FAIL, space_eval

    search_space = {'optimizer': hp.choice('optimizer', ['Adadelta', 'Adam']),
                    'learning_rate': hp.loguniform('learning_rate', -5, 0),
                    'dropout': hp.uniform('dropout', 0.1, 0.9)}                       # reference P2/01:194-198
    best = fmin(fn=objective, space=search_space, algo=tpe.suggest, max_evals=20,
                trials=ParallelTrials(parallelism=4))                                  # SparkTrials(parallelism=4)

* `Trials()` (default) evaluates sequentially in the driver process - required when every trial itself launches a
  distributed job through `Runner` (reference P2/02:342-365);
* `ParallelTrials(parallelism=k)` is `SparkTrials`: up to k single-GPU trials in flight (threads, each pinned to a
  GPU round-robin through `ParallelTrials.device_for`), suggestions drawn from whatever has completed so far.
* `fmin` returns the best point with `hp.choice` dimensions as INDICES (Hyperopt behaviour);
  `space_eval(space, best)` maps it back to values.  Exceptions in the objective become `STATUS_FAIL` trials.
* With an active tracking run, every trial is logged as a nested child run (params + loss), like SparkTrials +
  MLflow autologging (reference P2/01:217).
"""
from __future__ import annotations

import concurrent.futures as cf
import math
import os
import threading
import time
import traceback
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

STATUS_OK = "ok"
STATUS_FAIL = "fail"
STATUS_NEW = "new"
STATUS_RUNNING = "running"


# ------------------------------------------------------------------------------------------------ search space DSL
class _Node:
    def __init__(self, label: str, kind: str, **kw):
        self.label, self.kind, self.kw = label, kind, kw

    def __repr__(self):
        return f"hp.{self.kind}({self.label!r}, {self.kw})"


class _HP:
    @staticmethod
    def choice(label: str, options: Sequence[Any]) -> _Node:
        return _Node(label, "choice", options=list(options))

    @staticmethod
    def uniform(label: str, low: float, high: float) -> _Node:
        return _Node(label, "uniform", low=float(low), high=float(high))

    @staticmethod
    def loguniform(label: str, low: float, high: float) -> _Node:
        """exp(uniform(low, high)) - `hp.loguniform('learning_rate', -5, 0)` spans e^-5 .. 1."""
        return _Node(label, "loguniform", low=float(low), high=float(high))

    @staticmethod
    def quniform(label: str, low: float, high: float, q: float) -> _Node:
        return _Node(label, "quniform", low=float(low), high=float(high), q=float(q))

    @staticmethod
    def qloguniform(label: str, low: float, high: float, q: float) -> _Node:
        return _Node(label, "qloguniform", low=float(low), high=float(high), q=float(q))

    @staticmethod
    def normal(label: str, mu: float, sigma: float) -> _Node:
        return _Node(label, "normal", mu=float(mu), sigma=float(sigma))

    @staticmethod
    def randint(label: str, low: int, high: Optional[int] = None) -> _Node:
        if high is None:
            low, high = 0, low
        return _Node(label, "randint", low=int(low), high=int(high))


hp = _HP()


def _nodes(space) -> List[_Node]:
    out: List[_Node] = []

    def walk(s):
        if isinstance(s, _Node):
            out.append(s)
            if s.kind == "choice":
                for o in s.kw["options"]:
                    walk(o)
        elif isinstance(s, dict):
            for v in s.values():
                walk(v)
        elif isinstance(s, (list, tuple)):
            for v in s:
                walk(v)

    walk(space)
    seen, uniq = set(), []
    for n in out:
        if n.label not in seen:
            seen.add(n.label)
            uniq.append(n)
    return uniq


def space_eval(space, point: Dict[str, Any]):
    """Replace every hp node by the value `point` assigns to its label (choice: index -> option)."""
    if isinstance(space, _Node):
        v = point[space.label]
        if space.kind == "choice":
            return space_eval(space.kw["options"][int(v)], point)
        return v
    if isinstance(space, dict):
        return {k: space_eval(v, point) for k, v in space.items()}
    if isinstance(space, (list, tuple)):
        return type(space)(space_eval(v, point) for v in space)
    return space


def _sample_node(n: _Node, rng: np.random.Generator):
    k = n.kw
    if n.kind == "choice":
        return int(rng.integers(0, len(k["options"])))
    if n.kind == "uniform":
        return float(rng.uniform(k["low"], k["high"]))
    if n.kind == "loguniform":
        return float(math.exp(rng.uniform(k["low"], k["high"])))
    if n.kind == "quniform":
        return float(np.round(rng.uniform(k["low"], k["high"]) / k["q"]) * k["q"])
    if n.kind == "qloguniform":
        return float(np.round(math.exp(rng.uniform(k["low"], k["high"])) / k["q"]) * k["q"])
    if n.kind == "normal":
        return float(rng.normal(k["mu"], k["sigma"]))
    if n.kind == "randint":
        return int(rng.integers(k["low"], k["high"]))
    raise ValueError(n.kind)


# ------------------------------------------------------------------------------------------------ trials
class Trials:
    """Sequential trial store (Hyperopt's default `Trials`)."""

    parallelism = 1

    def __init__(self):
        self.trials: List[dict] = []
        self._lock = threading.Lock()

    def new_trial(self, vals: Dict[str, Any]) -> dict:
        with self._lock:
            t = {"tid": len(self.trials), "state": STATUS_NEW, "result": {"status": STATUS_NEW},
                 "misc": {"vals": dict(vals)}, "book_time": time.time(), "refresh_time": None}
            self.trials.append(t)
            return t

    @property
    def results(self) -> List[dict]:
        return [t["result"] for t in self.trials]

    def losses(self) -> List[Optional[float]]:
        return [t["result"].get("loss") if t["result"].get("status") == STATUS_OK else None for t in self.trials]

    def completed(self) -> List[dict]:
        return [t for t in self.trials if t["result"].get("status") == STATUS_OK]

    @property
    def best_trial(self) -> dict:
        done = self.completed()
        if not done:
            raise ValueError("no successful trial")
        return min(done, key=lambda t: t["result"]["loss"])

    @property
    def argmin(self) -> Dict[str, Any]:
        return dict(self.best_trial["misc"]["vals"])

    def __len__(self):
        return len(self.trials)


class ParallelTrials(Trials):
    """`SparkTrials(parallelism=k)`: k trials in flight (reference P2/01:226-238).

    ``executor='process'`` (the default on a GPU box, ``'auto'``): every slot is a persistent worker PROCESS pinned to
    GPU ``slot % device_count`` - its own CUDA context, its own CUDA-graph captures, its own extension state - that
    evaluates one trial at a time, like a Spark executor running one Hyperopt task; the objective travels by value
    (cloudpickle) once per worker.  ``executor='thread'`` evaluates in threads of the driver process (CPU objectives).
    `device_for(tid)` names the GPU a trial should use (inside a process worker: the worker's own device)."""

    def __init__(self, parallelism: int = 4, timeout: Optional[float] = None, executor: str = "auto"):
        super().__init__()
        self.parallelism = max(1, int(parallelism))
        self.timeout = timeout
        if executor not in ("auto", "thread", "process"):
            raise ValueError("executor must be 'auto', 'thread' or 'process'")
        self.executor = executor

    def uses_processes(self) -> bool:
        if self.executor == "auto":
            try:
                import torch

                return torch.cuda.is_available() and os.environ.get("B200DDL_FORCE_CPU", "0") != "1"
            except Exception:
                return False
        return self.executor == "process"

    @staticmethod
    def device_for(tid: int) -> int:
        import torch

        n = torch.cuda.device_count() if torch.cuda.is_available() else 0
        slot = os.environ.get("B200DDL_TRIAL_SLOT")
        if slot is not None:
            return int(slot) % n if n else -1
        return tid % n if n else -1


SparkTrials = ParallelTrials

_trial_ctx = threading.local()


def current_trial() -> Optional[dict]:
    """Inside an objective: the trial being evaluated (tid, misc.vals)."""
    return getattr(_trial_ctx, "trial", None)


# ------------------------------------------------------------------------------------------------ algorithms
class _Rand:
    @staticmethod
    def suggest(nodes: List[_Node], trials: Trials, rng: np.random.Generator) -> Dict[str, Any]:
        return {n.label: _sample_node(n, rng) for n in nodes}


class _TPE:
    """Tree-structured Parzen Estimator (Bergstra et al. 2011), independent per dimension: observations are split
    at the gamma-quantile of the loss into good/bad; candidates are drawn from the good density l(x) and the one
    maximising l(x)/g(x) is proposed."""

    n_startup_jobs = 20
    gamma = 0.25
    n_ei_candidates = 24
    prior_weight = 1.0

    @classmethod
    def suggest(cls, nodes, trials, rng):
        done = trials.completed()
        if len(done) < min(cls.n_startup_jobs, 5 * max(1, len(nodes))):
            return _Rand.suggest(nodes, trials, rng)
        losses = np.array([t["result"]["loss"] for t in done], dtype=np.float64)
        order = np.argsort(losses)
        n_good = max(1, int(math.ceil(cls.gamma * math.sqrt(len(done)) * 1.0)))
        n_good = min(max(n_good, int(math.ceil(cls.gamma * len(done)))), len(done) - 1) if len(done) > 1 else 1
        good_idx, bad_idx = order[:n_good], order[n_good:]
        out = {}
        for n in nodes:
            obs = [t["misc"]["vals"].get(n.label) for t in done]
            good = [obs[i] for i in good_idx if obs[i] is not None]
            bad = [obs[i] for i in bad_idx if obs[i] is not None]
            out[n.label] = cls._suggest_dim(n, good, bad, rng)
        return out

    @classmethod
    def _suggest_dim(cls, n: _Node, good, bad, rng):
        k = n.kw
        if n.kind in ("choice", "randint"):
            lo = 0 if n.kind == "choice" else k["low"]
            m = len(k["options"]) if n.kind == "choice" else k["high"] - k["low"]

            def probs(vals):
                c = np.full(m, cls.prior_weight / m)
                for v in vals:
                    c[int(v) - lo] += 1.0
                return c / c.sum()

            pl, pg = probs(good), probs(bad)
            cand = rng.choice(m, size=cls.n_ei_candidates, p=pl)
            best = cand[np.argmax(np.log(pl[cand]) - np.log(pg[cand]))]
            return int(best) + lo
        # continuous: work in the transformed (uniform / log / identity) space
        log = n.kind in ("loguniform", "qloguniform")
        if n.kind == "normal":
            lo_, hi_ = k["mu"] - 4 * k["sigma"], k["mu"] + 4 * k["sigma"]
        else:
            lo_, hi_ = k["low"], k["high"]
        tf = (lambda v: math.log(max(v, 1e-300))) if log else (lambda v: float(v))
        g = np.array([tf(v) for v in good], dtype=np.float64)
        b = np.array([tf(v) for v in bad], dtype=np.float64)

        def parzen(obs):
            mus = np.concatenate([obs, [(lo_ + hi_) / 2.0]])  # prior component in the middle
            srt = np.sort(mus)
            if len(srt) > 1:
                gaps = np.diff(srt)
                sig_sorted = np.maximum(np.concatenate([[gaps[0]], np.maximum(gaps[:-1], gaps[1:]), [gaps[-1]]]), 1e-12)
                sig = np.empty_like(mus)
                sig[np.argsort(mus)] = sig_sorted
            else:
                sig = np.array([hi_ - lo_])
            sig = np.clip(sig, (hi_ - lo_) / min(100.0, 1.0 + len(mus)), hi_ - lo_)
            sig[-1] = hi_ - lo_
            w = np.ones(len(mus))
            w[-1] = cls.prior_weight
            return mus, sig, w / w.sum()

        def logpdf(x, mus, sig, w):
            z = (x[:, None] - mus[None, :]) / sig[None, :]
            comp = -0.5 * z * z - np.log(sig[None, :]) - 0.5 * math.log(2 * math.pi) + np.log(w[None, :])
            mx = comp.max(axis=1, keepdims=True)
            return (mx + np.log(np.exp(comp - mx).sum(axis=1, keepdims=True)))[:, 0]

        ml, sl, wl = parzen(g)
        mg, sg, wg = parzen(b)
        comp = rng.choice(len(ml), size=cls.n_ei_candidates, p=wl)
        cand = np.clip(rng.normal(ml[comp], sl[comp]), lo_, hi_)
        score = logpdf(cand, ml, sl, wl) - logpdf(cand, mg, sg, wg)
        x = float(cand[int(np.argmax(score))])
        v = math.exp(x) if log else x
        if n.kind in ("quniform", "qloguniform"):
            v = float(np.round(v / k["q"]) * k["q"])
        return v


class _AlgoModule:
    def __init__(self, impl):
        self._impl = impl

    def suggest(self, nodes, trials, rng):
        return self._impl.suggest(nodes, trials, rng)


tpe = _AlgoModule(_TPE)
rand = _AlgoModule(_Rand)


# ------------------------------------------------------------------------------------------------ process workers
def _trial_worker(slot: int, conn) -> None:
    """Persistent trial process: pinned to one GPU, receives the objective once, then (params, tracking context) per trial."""
    os.environ["B200DDL_TRIAL_SLOT"] = str(slot)
    try:
        import torch

        if torch.cuda.is_available() and os.environ.get("B200DDL_FORCE_CPU", "0") != "1":
            torch.cuda.set_device(slot % torch.cuda.device_count())
    except Exception:
        pass
    import cloudpickle

    from .. import tracking

    fn = None
    while True:
        try:
            msg = conn.recv()
        except EOFError:
            return
        if msg is None:
            return
        if msg[0] == "fn":
            fn = cloudpickle.loads(msg[1])
            continue
        _, params, uri, run_id, trial = msg
        opened = False
        try:
            if uri:
                tracking.set_tracking_uri(uri)
            if run_id is not None:
                tracking.start_run(run_id=run_id)  # re-open the trial's child run: the objective's logging lands there
                opened = True
            _trial_ctx.trial = trial
            res = fn(params)
            conn.send(("ok", cloudpickle.dumps(res)))
        except BaseException as e:  # the driver decides (catch_eval_exceptions) what a failed objective means
            conn.send(("err", repr(e), traceback.format_exc()))
        finally:
            _trial_ctx.trial = None
            if opened:
                try:
                    tracking._stack().pop()  # detach without ending: the driver ends the run
                except Exception:
                    pass


class _TrialProcessPool:
    """`n` trial workers started as `python -m b200ddl.hpo._trial_worker` (utils/procpool.py explains why not mp.spawn)."""

    def __init__(self, n: int, fn: Callable):
        import queue

        import cloudpickle

        from ..utils.procpool import start_worker

        blob = cloudpickle.dumps(fn)
        self.procs, self.conns = [], []
        self.idle: "queue.Queue[int]" = queue.Queue()
        for i in range(n):
            p, c = start_worker("b200ddl.hpo._trial_worker", i)
            c.send(("fn", blob))
            self.procs.append(p)
            self.conns.append(c)
            self.idle.put(i)

    def call(self, params, trial: dict):
        import cloudpickle

        from .. import tracking

        i = self.idle.get()
        try:
            active = tracking.active_run()
            light = {"tid": trial["tid"], "misc": trial["misc"]}
            self.conns[i].send(("eval", params, tracking.get_tracking_uri(), active.info.run_id if active else None, light))
            try:
                msg = self.conns[i].recv()
            except EOFError:
                raise RuntimeError(f"trial worker {i} died (exit code {self.procs[i].poll()})") from None
            if msg[0] == "ok":
                return cloudpickle.loads(msg[1])
            raise RuntimeError(f"objective failed in trial worker {i}: {msg[1]}\n{msg[2]}")
        finally:
            self.idle.put(i)

    def close(self) -> None:
        for c in self.conns:
            try:
                c.send(None)
                c.close()
            except Exception:
                pass
        for p in self.procs:
            try:
                p.wait(timeout=10)
            except Exception:
                p.kill()


# ------------------------------------------------------------------------------------------------ fmin
def _run_trial(fn: Callable, space, trial: dict, catch: bool) -> None:
    from .. import tracking

    trial["state"] = STATUS_RUNNING
    params = space_eval(space, trial["misc"]["vals"])
    _trial_ctx.trial = trial
    parent = getattr(_trial_ctx, "parent_run_id", None)
    child = None
    try:
        if parent is not None:
            # a worker thread has no active-run stack of its own: re-open the parent, then nest
            if tracking.active_run() is None:
                tracking.start_run(run_id=parent)
                _trial_ctx.opened_parent = True
            child = tracking.start_run(run_name=f"trial_{trial['tid']}", nested=True)
            flat = params if isinstance(params, dict) else {"params": params}
            tracking.log_params({k: v for k, v in flat.items() if isinstance(v, (int, float, str, bool))})
        t0 = time.time()
        try:
            res = fn(params)
            if not isinstance(res, dict):
                res = {"loss": float(res), "status": STATUS_OK}
            if res.get("status") == STATUS_OK:
                res["loss"] = float(res["loss"])
        except Exception as e:  # objectives may fail: record and continue
            if not catch:
                raise
            res = {"status": STATUS_FAIL, "error": repr(e), "traceback": traceback.format_exc()}
        res["eval_time"] = time.time() - t0
        trial["result"] = res
        trial["state"] = res.get("status", STATUS_FAIL)
        trial["refresh_time"] = time.time()
        if child is not None:
            if res.get("status") == STATUS_OK:
                tracking.log_metric("loss", res["loss"])
            tracking.set_tag("hpo.status", res.get("status"))
    finally:
        if child is not None:
            tracking.end_run()
        if getattr(_trial_ctx, "opened_parent", False):
            tracking._stack().pop()  # detach without ending the driver's run
            _trial_ctx.opened_parent = False
        _trial_ctx.trial = None


def fmin(fn: Callable[[Any], Any], space, algo=None, max_evals: int = 10, trials: Optional[Trials] = None,
         rstate=None, catch_eval_exceptions: bool = True, verbose: bool = False, show_progressbar: bool = False,
         return_argmin: bool = True, timeout: Optional[float] = None, loss_threshold: Optional[float] = None,
         early_stop_fn: Optional[Callable] = None, **hyperopt_kwargs):
    """Minimise `fn(params) -> loss | {'loss', 'status'}` over `space` (reference P2/01:229-238, P2/02:360-365).

    Hyperopt's stopping arguments are honoured: `timeout` (seconds; no new trial starts after it), `loss_threshold` (stop
    once the best loss is at or below it) and `early_stop_fn(trials, *state) -> (stop, state)`.  Other Hyperopt keyword
    arguments (`max_queue_len`, `points_to_evaluate`, ...) are accepted and ignored with a warning."""
    from .. import tracking

    if hyperopt_kwargs:
        import warnings

        warnings.warn(f"hpo.fmin ignores {sorted(hyperopt_kwargs)}", stacklevel=2)
    t_start = time.time()
    stop_state: list = []

    def should_stop() -> bool:
        nonlocal stop_state
        if timeout is not None and time.time() - t_start >= timeout:
            return True
        done = trials.completed()
        if loss_threshold is not None and done and min(t["result"]["loss"] for t in done) <= loss_threshold:
            return True
        if early_stop_fn is not None and len(trials):
            stop, stop_state = early_stop_fn(trials, *stop_state)
            return bool(stop)
        return False

    algo = algo or tpe
    suggest = algo.suggest if hasattr(algo, "suggest") else algo
    if trials is None:
        trials = Trials()
    if isinstance(rstate, np.random.Generator):
        rng = rstate
    else:
        rng = np.random.default_rng(rstate)
    nodes = _nodes(space)
    active = tracking.active_run()
    parent_id = active.info.run_id if active is not None else None
    n_par = getattr(trials, "parallelism", 1)

    pool = None
    call = fn
    if n_par > 1 and hasattr(trials, "uses_processes") and trials.uses_processes():
        pool = _TrialProcessPool(n_par, fn)

        def call(params):  # runs in a driver thread; the objective itself runs in a pinned worker process
            return pool.call(params, current_trial())

    def evaluate(trial):
        _trial_ctx.parent_run_id = parent_id
        _run_trial(call, space, trial, catch_eval_exceptions)
        if verbose:
            r = trial["result"]
            print(f"[hpo] trial {trial['tid']} {r.get('status')} loss={r.get('loss')} vals={trial['misc']['vals']}")

    if n_par <= 1:
        while len(trials) < max_evals and not should_stop():
            evaluate(trials.new_trial(suggest(nodes, trials, rng)))
    else:
        deadline = (time.time() + trials.timeout) if getattr(trials, "timeout", None) else None
        with cf.ThreadPoolExecutor(n_par) as ex:
            pending = set()
            stopping = False
            while (len(trials) < max_evals and not stopping) or pending:
                stopping = stopping or should_stop()      # running trials finish; no new ones start
                while len(trials) < max_evals and len(pending) < n_par and not stopping:
                    pending.add(ex.submit(evaluate, trials.new_trial(suggest(nodes, trials, rng))))
                if not pending:
                    break
                done, pending = cf.wait(pending, return_when=cf.FIRST_COMPLETED,
                                        timeout=None if deadline is None else max(0.0, deadline - time.time()))
                for d in done:
                    d.result()
                if deadline is not None and time.time() > deadline:
                    break
        if pool is not None:
            pool.close()
    if not trials.completed():
        raise RuntimeError("all trials failed:\n" + "\n".join(str(t["result"].get("error")) for t in trials.trials))
    return trials.argmin if return_argmin else trials


__all__ = ["fmin", "hp", "tpe", "rand", "Trials", "ParallelTrials", "SparkTrials", "STATUS_OK", "STATUS_FAIL",
           "space_eval", "current_trial"]
