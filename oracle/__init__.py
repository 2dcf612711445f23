"""CPU oracle for the DeepVCP registration hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package, and only as
the checker. The product package (``deepvcp-pointcloud-registration_b200``)
never imports it and has no CPU fallback.

Contents
  dvcp_oracle.c       C restatement of FPS, ball query, KNN, candidate grid.
  native.py           ctypes binding of the above.
  stages.py           torch-fp32/fp64 restatement of every stage and of the
                      whole ``DeepVCP.forward`` + pose solve.
  reference_shims.py  imports the unmodified reference from /root/reference
                      under the four shims of SURVEY App. C (build container
                      only; used to generate and re-check tests/golden/).

Pinning: the reference ships no tests or golden vectors. The oracle
is pinned against outputs of the reference itself, generated in the build
container by tests/golden/make_golden.py and committed under tests/golden/.
One stage is NOT pinnable that way: the K-nearest-neighbour search is the
third-party ``knn_cuda`` extension (github unlimblue/KNN_CUDA, unpinned, absent
from /root/reference and not installable offline). ``orc_knn_f32`` restates its
published algorithm and the reference's call sites; KNN parity is therefore
"parity unpinned" and every fixture that passes through KNN inherits that.
"""
