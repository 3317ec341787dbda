"""Draw output: the parquet dumps the reference's experiment scripts intended but left commented out
(``experiment_ar1.py:93-94``, ``experiment_funnel.py:58-59``: one column per ``model.parameter_names()``
entry), extended with ``chain`` and ``iteration`` columns for a batch of chains."""
from __future__ import annotations

import numpy as np


def draws_table(draws, names, chains=None, thin=1, first_iteration=0):
    """``draws``: (M, B, D) tensor / array or (M, D) array.  Returns a pyarrow Table in long format with columns
    ``chain``, ``iteration`` and one float64 column per parameter; ``chains`` keeps the first that many chains."""
    import pyarrow as pa
    arr = draws.detach().cpu().numpy() if hasattr(draws, "detach") else np.asarray(draws)
    if arr.ndim == 2:
        arr = arr[:, None, :]
    M, B, D = arr.shape
    if len(names) != D:
        raise ValueError(f"{len(names)} names for {D} parameters")
    if chains is not None:
        B = min(B, int(chains))
        arr = arr[:, :B]
    it = first_iteration + thin * np.arange(M, dtype=np.int64)
    cols = {"chain": np.tile(np.arange(B, dtype=np.int32), M), "iteration": np.repeat(it, B)}
    flat = np.ascontiguousarray(arr, dtype=np.float64).reshape(M * B, D)
    for k, name in enumerate(names):
        cols[name] = flat[:, k]
    return pa.table(cols)


def write_draws(path, draws, names, chains=None, thin=1, first_iteration=0):
    """Write ``draws`` to ``path`` (parquet).  Returns the number of rows written."""
    import pyarrow.parquet as pq
    t = draws_table(draws, names, chains=chains, thin=thin, first_iteration=first_iteration)
    pq.write_table(t, str(path))
    return t.num_rows
