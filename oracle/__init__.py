"""CPU oracle for the DiffBindFR reverse-diffusion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diffbindfr_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
as the CPU baseline that is timed beside the CUDA path.

Layout
------
``oracle/thirdparty/``  restatement of the published algorithms of the wheels the
    reference depends on but which are absent here (e3nn 0.5.1, torch-scatter 2.1.0,
    torch-cluster 1.6.0).
``oracle/shims/``       import shims that expose ``oracle/thirdparty`` under the
    third-party module names, so that the reference's own files run unmodified
    (oracle "O1"; only usable where ``/root/reference`` exists).
``oracle/model.py``, ``oracle/sampler.py``, ``oracle/geometry.py``
    clean-room restatement of the reference's own files (oracle "O2"), travels to
    the GPU box.

Parity pinning: the reference ships no tests or golden vectors,
so the pins are (1) O1 == O2 on seeded inputs, with the fixtures produced by
``tools/make_golden.py`` committed under ``tests/golden/``, (2) known-answer
anchors for the Clebsch-Gordan tensors / spherical harmonics, and (3) SE(3)
equivariance properties.  What cannot be pinned here is agreement of
``oracle/thirdparty`` with real e3nn / torch-cluster binaries (not installable,
no network): stated as residual risk in DESIGN.md.
"""
