"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's algorithms for the restoration hot path.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker / reported baseline.  The product
path (msd_b200 + libmsd_b200.so) never imports it and has no CPU fallback.

Pinning: the reference (ajammoussi/Multi-Spectral-Satellite-Image-Denoising) ships NO tests or golden vectors
(SURVEY 4).  Each module here is therefore pinned against the reference itself, imported from /root/reference in the
authoring container by ``oracle/make_golden.py``; the resulting vectors are committed under ``tests/golden/`` and the
CPU test-suite checks the oracle against them.  Parts with no executable reference counterpart are marked
"parity unpinned" in their docstrings (the timm-Block LayerNorm-eps variant; the CUDA Philox stream, which is pinned
on the GPU box against torch.randn / torch.rand themselves instead).
"""
