"""CPU oracle for the KLHR hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from this package.  The product
(``klhr_b200``) never imports it and has no CPU fallback.

Contents
--------
stan_models.py  NumPy fp64 restatement of the Stan programs the north star names
                (reference ``stan/{normal,ill-normal,funnel,corr-normal,ar1,arK,
                rosenbrock}.stan``), BridgeStan conventions propto=True, jacobian=True.
bsmodel.py      ``BSModel`` shim with the surface of reference ``bsmodel.py:5-55``
                backed by stan_models (BridgeStan itself is not installable here).
adapt.py        Restatement of ``onlinemoments.py``, ``onlinepca.py``,
                ``windowedadaptation.py`` plus the pooled (raw-sum) forms.
ref_port.py     Single-chain port of ``klhr.py`` / ``klhr_sinh.py`` that still calls
                SciPy BFGS exactly like the reference (``klhr.py:126-141``); pinned
                BIT-EXACT against tapes of the unmodified reference
                (``tests/golden/*.npz``, made by ``oracle/make_golden.py``).
batched.py      Batched restatement of the same step with the fixed-iteration Newton
                optimiser the CUDA kernels use (only ``minimize`` is replaced);
                pinned against the same tapes at optimiser tolerance.
philox.py       NumPy Philox4x32-10 used to check the in-kernel RNG streams.
make_golden.py  Runs the UNMODIFIED reference from /root/reference under a tape RNG
                (authoring container only) and writes the fixtures.

Parity status: the sampler arithmetic is pinned against the live reference; the
model layer (BridgeStan) and SciPy are third-party, un-vendored and un-pinned in the
reference -- see DESIGN.md "Oracle".
"""
