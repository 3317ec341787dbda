"""Development aid: which host-side pattern hangs under `ncu` (observed: NumPy LAPACK after a fork once SciPy's BLAS is loaded)."""
import sys, time, subprocess
sys.path.insert(0, '.')
mode = sys.argv[1]
import numpy as np
if "scipy" in mode:
    import scipy.linalg
if "fork" in mode:
    p = subprocess.Popen(["sleep", "1"])
if "spawn" in mode:
    p = subprocess.Popen(["/bin/sleep", "1"], close_fds=False)
N = 256
idx = np.arange(N); Sigma = 0.9 ** np.abs(idx[:, None] - idx[None, :])
t = time.time(); P = np.linalg.inv(Sigma); print(mode, 'inv ok', round(time.time() - t, 4), flush=True)
