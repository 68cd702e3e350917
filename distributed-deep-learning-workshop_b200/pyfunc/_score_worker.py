"""Entry point of a `spark_udf` scoring worker: `python -m b200ddl.pyfunc._score_worker <address> <slot>`."""
from ..utils.procpool import connect_parent
from . import _pool_worker

if __name__ == "__main__":
    conn, slot = connect_parent()
    _pool_worker(slot, conn)
