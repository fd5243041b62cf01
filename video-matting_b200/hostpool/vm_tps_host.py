"""Host side of the thin-plate-spline path (reference tps.py:78-119): the (N+3) x (N+3) system and numpy's
own truncated pseudo-inverse.  NumPy only - no torch, no relative imports - because `SolverPool` runs these
functions in spawned worker processes that import this file by its bare module name (the directory of this
file is put on sys.path; nothing else of the package is importable from there).

The solve is *called*, not re-implemented: at 1080p / 4K the rcond=1e-15 truncation of `np.linalg.pinv` is
active, so only the same LAPACK call reproduces the reference's coefficients bit for bit.
"""
import os

import numpy as np


def tps_kernel_matrix(points):
    """L = [[K, P], [P^T, 0]], K_ab = U(|P_a - P_b|), U(r) = r^2 log r (reference tps.py:78-98).
    The expression order is the reference's so that L - and therefore numpy's truncated
    pseudo-inverse of it - is bit-identical."""
    pts = np.asarray(points, dtype=np.float64)
    n = len(pts)
    d0 = np.subtract.outer(pts[:, 0], pts[:, 0])
    d1 = np.subtract.outer(pts[:, 1], pts[:, 1])
    r = np.sqrt(d0 ** 2 + d1 ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        K = (r ** 2) * np.where(r < 1e-100, 0, np.log(r))
    L = np.zeros((n + 3, n + 3))
    L[:n, :n] = K
    L[:n, n] = 1.0
    L[:n, n + 1:] = pts
    L[n:, :n] = L[:n, n:].T
    return L


def tps_solve(src_points, dst_points):
    """Spline coefficients (N+3, 2) mapping src_points onto dst_points: numpy's own
    ``dot(pinv(L), V)`` (reference tps.py:113-119)."""
    dst = np.asarray(dst_points, dtype=np.float64)
    V = np.zeros((len(dst) + 3, 2))
    V[:len(dst)] = dst
    return np.dot(np.linalg.pinv(tps_kernel_matrix(src_points)), V)


_stacked_kernel_ok = [None]      # None: not checked yet in this process


def tps_kernel_matrices(points):
    """tps_kernel_matrix for a stack (m, N, 2) of control-point sets with the same element-wise expressions.
    numpy's sqrt / log loops are expected to give the same value for an element wherever it sits in an array; that
    is verified once per process on the first stack (bit for bit against the per-frame function), and the per-frame
    function is used from then on if it ever fails."""
    pts = np.asarray(points, dtype=np.float64)
    m, n = pts.shape[:2]
    d0 = pts[:, :, None, 0] - pts[:, None, :, 0]
    d1 = pts[:, :, None, 1] - pts[:, None, :, 1]
    r = np.sqrt(d0 ** 2 + d1 ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        K = (r ** 2) * np.where(r < 1e-100, 0, np.log(r))
    L = np.zeros((m, n + 3, n + 3))
    L[:, :n, :n] = K
    L[:, :n, n] = 1.0
    L[:, :n, n + 1:] = pts
    L[:, n:, :n] = np.transpose(L[:, :n, n:], (0, 2, 1))
    return L


def _kernel_stack(chunk):
    if _stacked_kernel_ok[0] is not False:
        L = tps_kernel_matrices(np.stack([np.asarray(d, dtype=np.float64) for (_, d) in chunk]))
        if _stacked_kernel_ok[0] is None:
            ref = np.stack([tps_kernel_matrix(d) for (_, d) in chunk])
            _stacked_kernel_ok[0] = bool(L.tobytes() == ref.tobytes())
            return ref
        return L
    return np.stack([tps_kernel_matrix(d) for (_, d) in chunk])


def solve_chunk(chunk):
    """np.dot(np.linalg.pinv(L), V) for a list of (grid, deformed grid) pairs: numpy's stacked pinv runs
    the same LAPACK call per matrix as the reference's per-frame call (bit-identical, checked in the tests)."""
    L = _kernel_stack(chunk)
    V = np.zeros((len(chunk), L.shape[1], 2))
    for k, (g, _) in enumerate(chunk):
        g = np.asarray(g, dtype=np.float64)
        V[k, :len(g)] = g
    Li = np.linalg.pinv(L)
    return np.stack([np.dot(Li[k], V[k]) for k in range(len(chunk))])


def solve_many(grids, stack=16):
    """(ctrl (n, N, 2), coef (n, N+3, 2)) for a list of (regular grid, deformed grid) pairs as used by
    augmentation.warp_image(..., thin=grids): the system is built from the DEFORMED grid and maps back onto
    the regular one (reference tps.py:51)."""
    grids = list(grids)
    ctrl = np.stack([np.asarray(d, dtype=np.float64) for (_, d) in grids])
    if len({len(d) for (_, d) in grids}) == 1:
        coef = np.concatenate([solve_chunk(grids[i:i + stack]) for i in range(0, len(grids), stack)])
    else:
        coef = np.stack([tps_solve(d, g) for (g, d) in grids])
    return ctrl, coef


# ---------------------------------------------------------------------------------------------------------
# worker processes
# ---------------------------------------------------------------------------------------------------------

def _worker_main():
    """`python vm_tps_host.py --worker`: length-prefixed pickles in on stdin ((id, grids)), out on stdout
    ((id, coef) or (id, exception)).  Plain subprocesses - not multiprocessing - so that nothing depends on how
    the parent's __main__ is written (spawned multiprocessing children re-import it)."""
    import pickle
    import struct
    import sys
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)                  # tiny systems: the workers are the parallelism
    except Exception:
        pass
    rd, wr = sys.stdin.buffer, sys.stdout.buffer
    while True:
        head = rd.read(8)
        if len(head) < 8:
            return
        (n,) = struct.unpack("<q", head)
        tid, grids = pickle.loads(rd.read(n))
        try:
            res = (tid, solve_many(grids)[1])
        except Exception as e:               # noqa: BLE001 - handed to the caller
            res = (tid, e)
        blob = pickle.dumps(res, protocol=pickle.HIGHEST_PROTOCOL)
        wr.write(struct.pack("<q", len(blob)))
        wr.write(blob)
        wr.flush()


class SolverPool:
    """`workers` processes that run `solve_many` on slices of a clip's grids, so that the host solve of clip k+1
    overlaps the kernels of clip k and scales with host cores (np.linalg.pinv holds the GIL, threads do not
    help).  Results are bit-identical to the in-process call: same function, same numpy, one BLAS thread."""

    def __init__(self, workers):
        import subprocess
        import sys
        import threading
        self.workers = max(1, int(workers))
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        self._procs, self._readers = [], []
        self._results, self._cv = {}, threading.Condition()
        self._wlock = [threading.Lock() for _ in range(self.workers)]
        self._next_id, self._rr = 0, 0
        for k in range(self.workers):
            p = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker"], stdin=subprocess.PIPE,
                                 stdout=subprocess.PIPE, env=env)
            t = threading.Thread(target=self._drain, args=(p,), daemon=True)
            t.start()
            self._procs.append(p)
            self._readers.append(t)

    def _drain(self, p):
        import pickle
        import struct
        while True:
            head = p.stdout.read(8)
            if len(head) < 8:
                with self._cv:
                    self._results[("dead", p.pid)] = True
                    self._cv.notify_all()
                return
            (n,) = struct.unpack("<q", head)
            tid, val = pickle.loads(p.stdout.read(n))
            with self._cv:
                self._results[tid] = val
                self._cv.notify_all()

    def submit(self, grids, per_task=16):
        """Start solving; returns a handle for `collect` (handles may be collected in any order)."""
        import pickle
        import struct
        grids = [(np.asarray(g, dtype=np.float64), np.asarray(d, dtype=np.float64)) for (g, d) in grids]
        ctrl = np.stack([d for (_, d) in grids])
        ids = []
        for i in range(0, len(grids), per_task):
            tid, self._next_id = self._next_id, self._next_id + 1
            k, self._rr = self._rr, (self._rr + 1) % self.workers
            blob = pickle.dumps((tid, grids[i:i + per_task]), protocol=pickle.HIGHEST_PROTOCOL)
            with self._wlock[k]:
                self._procs[k].stdin.write(struct.pack("<q", len(blob)))
                self._procs[k].stdin.write(blob)
                self._procs[k].stdin.flush()
            ids.append(tid)
        return ctrl, ids

    def collect(self, handle, timeout=120.0):
        ctrl, ids = handle
        parts = []
        with self._cv:
            for tid in ids:
                while tid not in self._results:
                    if any(p.poll() is not None for p in self._procs):
                        raise RuntimeError("SolverPool: a worker process died")
                    if not self._cv.wait(timeout):
                        raise TimeoutError("SolverPool: no result within %.0f s" % timeout)
                val = self._results.pop(tid)
                if isinstance(val, Exception):
                    raise val
                parts.append(val)
        return ctrl, np.concatenate(parts)

    def solve(self, grids, per_task=16):
        return self.collect(self.submit(grids, per_task))

    def close(self):
        for p in self._procs:
            try:
                p.stdin.close()
            except Exception:
                pass
        for p in self._procs:
            try:
                p.wait(timeout=5)
            except Exception:
                p.kill()
        self._procs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


if __name__ == "__main__":
    import sys
    if "--worker" in sys.argv:
        _worker_main()
