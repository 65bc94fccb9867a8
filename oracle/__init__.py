"""oracle/ — TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).

CPU restatement (plain PyTorch fp32/fp64 on the host) of the MM-DTI training hot path:
Gaussian pair-distance basis -> pair bias -> pair-biased self-attention encoder,
InfoNCE, SupCon/ConR, FDS.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there only
as the checker or the timed CPU baseline.

Pinning status (see DESIGN.md "Oracle"):
  * Everything that lives in the reference's own files (models/transformers.py,
    models/mm_model.py gaussian/GaussianLayer/NonLinearHead, models/infonce.py,
    models/contrastive.py, models/fds.py, utils/util.py:calibrate_mean_var) is PINNED:
    ``oracle/make_golden.py`` executes those very files from /root/reference (read-only)
    and ``tests/golden/*.npz`` holds their outputs on seeded inputs; tests compare
    ``oracle.restate`` to those fixtures.
  * The attention-layer arithmetic lives in third-party Uni-Core
    (github.com/dptech-corp/Uni-Core, NOT vendored, NOT pinned by the reference, absent
    here): ``oracle/shims/unicore`` restates its published algorithm.  For that part the
    parity is UNPINNED (no reference-side golden vector exists).
"""
