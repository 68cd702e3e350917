"""Where does CPU JPEG decode stop scaling on this host?  (no GPU needed; run on the GPU box because its CPU is what feeds it)

1. host facts: logical CPUs, affinity mask, cgroup CPU quota, CPU model, load average
2. one process, one thread: PIL decode + resize ms per image (the unit cost)
3. N bare decode processes (loader/_decode_worker.py), one driver thread each, no ring, no parquet: images/s for N = 8..96
   -> the host's ceiling for process-parallel decode, independent of the loader's design
4. the loader itself (`make_dataset(decode_processes=P, workers_count=T)`, device='cpu' ring): images/s for a few (T, P)

Lines are prefixed CPU_DECODE.
"""
import json
import os
import sys
import tempfile
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from b200ddl import Session
from b200ddl.data import col, pandas_udf, synthetic_images
from b200ddl.loader import make_converter
from b200ddl.models import decode_image
from b200ddl.utils.procpool import start_script_workers

H = W = 224


def host_facts():
    facts = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "loadavg": os.getloadavg()}
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
        try:
            facts[p] = open(p).read().strip()
        except OSError:
            pass
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                facts["model"] = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return facts


def bare_processes(blobs, n_procs, seconds=3.0, chunk=16):
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "distributed-deep-learning-workshop_b200",
                          "loader", "_decode_worker.py")
    t0 = time.time()
    procs = start_script_workers(script, n_procs, env={"OMP_NUM_THREADS": "1"})
    start_s = time.time() - t0
    head = np.array([H, W, chunk] + [len(b) for b in blobs[:chunk]], dtype=np.int32).tobytes()
    frame = b"".join([head] + blobs[:chunk])
    counts = [0] * n_procs
    stop = time.time() + seconds
    buf = [bytearray(chunk * H * W * 3) for _ in range(n_procs)]

    def drive(k):
        conn = procs[k][1]
        while time.time() < stop:
            conn.send_bytes(frame)
            conn.recv_bytes_into(buf[k])
            counts[k] += chunk

    ths = [threading.Thread(target=drive, args=(k,)) for k in range(n_procs)]
    t0 = time.time()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.time() - t0
    for p, c in procs:
        try:
            c.send_bytes(b"")
            c.close()
        except Exception:
            pass
    for p, _ in procs:
        try:
            p.wait(timeout=5)
        except Exception:
            p.kill()
    return sum(counts) / dt, start_s


def main():
    print("CPU_DECODE " + json.dumps({"host": host_facts()}), flush=True)
    root = tempfile.mkdtemp(prefix="b200ddl_cpudec_")
    Session(user="cpu@example.com", root=root)
    raw = synthetic_images(1024, size=(375, 500), jpeg=True, seed=11)
    blobs = [bytes(b) for b in raw.limit(64).to_pandas()["content"]]
    t0 = time.perf_counter()
    for i in range(128):
        decode_image(blobs[i % 64], (H, W))
    unit = (time.perf_counter() - t0) / 128 * 1e3
    print("CPU_DECODE " + json.dumps({"single_thread_ms_per_image": round(unit, 3), "jpeg_kb": round(len(blobs[0]) / 1024, 1)}), flush=True)
    for n in (8, 32, 64, 96, 120) if (os.cpu_count() or 1) < 32 else (32, 64, 96, 120):
        if n > (os.cpu_count() or 1):
            break
        rate, start_s = bare_processes(blobs, n)
        print("CPU_DECODE " + json.dumps({"bare_processes": n, "images_per_sec": round(rate, 1), "per_process": round(rate / n, 1),
                                          "ideal": round(n * 1e3 / unit, 1), "start_s": round(start_s, 2)}), flush=True)

    @pandas_udf("int")
    def label_idx(path):
        return path.map(lambda p: hash(p.split("/")[-2]) % 5)

    table = raw.withColumn("label_idx", label_idx(col("path"))).select(["content", "label_idx"])
    conv = make_converter(table, os.path.join(root, "cache"))
    for threads, procs in ((2, 8), (2, 48), (4, 96), (4, 120), (32, 0)):
        if procs > (os.cpu_count() or 1):
            continue
        with conv.make_dataset(batch_size=256, num_epochs=None, workers_count=threads, image_size=(H, W), device="cpu",
                               decode_processes=procs) as ds:
            it = iter(ds)
            for _ in range(6):
                next(it)
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < 4.0:
                next(it)
                n += 256
            rate = n / (time.perf_counter() - t0)
        print("CPU_DECODE " + json.dumps({"loader_threads": threads, "decode_processes": procs, "images_per_sec": round(rate, 1)}),
              flush=True)
        time.sleep(1.5)   # decode threads of the closed dataset finish the batch they were in
    conv.delete()


if __name__ == "__main__":
    main()
