"""Entry point of a `ParallelTrials(executor='process')` worker: `python -m b200ddl.hpo._trial_worker <address> <slot>`."""
from ..utils.procpool import connect_parent
from . import _trial_worker

if __name__ == "__main__":
    conn, slot = connect_parent()
    _trial_worker(slot, conn)
