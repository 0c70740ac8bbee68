"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the DOSTransformer hot path used as the
parity checker for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and only as the checker / the CPU baseline -- never as the thing that
is shipped or measured as the product.  ``dostransformer_b200`` never imports it.

Parity pin: the reference (``/root/reference``) ships no tests or golden vectors
(SURVEY.md section 4).  The oracle is therefore pinned against outputs of the
reference's own model code run in the build container (``oracle/make_golden.py``
imports ``/root/reference`` unchanged behind the stand-ins in ``oracle/shims.py``
and writes ``tests/golden/*.pt``).  The five third-party primitives the reference
calls (torch_scatter.scatter_sum/mean, torch_geometric.utils.to_dense_batch,
e3nn spherical_harmonics(l<=1)/smooth_cutoff) are NOT vendored and not installed
anywhere we can reach: for those "parity is unpinned" by the reference and the
published semantics restated in ``oracle/shims.py`` are the spec.
"""
