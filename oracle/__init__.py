"""CPU oracle for the InstaGeo chip-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU arm -- never as a fallback for the CUDA path.

What it restates (reference files are cited per function):

* ``oracle.preprocess``  -- numpy: band gather, ``* constant_multiplier``, nodata
  mask, f64->f32 round trip, per-band normalise, ``[C,T,H,W]`` permute, Fmask bit
  decode / cloud masking, strided window grid (``instageo/model/dataloader.py``,
  ``instageo/data/hls_utils.py``, ``instageo/data/data_pipeline.py``).
* ``oracle.prithvi``     -- torch CPU fp32: ``PrithviViT`` + timm ``Block`` +
  ``PrithviSeg`` head (``instageo/model/pritvhi.py``, ``instageo/model/model.py``;
  the transformer block arithmetic lives in the third-party dependency
  **timm==1.0.20**, absent from the reference tree and from this image, restated
  from its published definition).
* ``oracle.stitch``      -- numpy: overlap-averaging stitch of sliding-window
  logits (OUR specification, SURVEY.md Appendix A.6 -- the reference snapshot
  ships no stitch; parity for the averaging itself is therefore "unpinned").

Pinning status (see DESIGN.md "Oracle"):

* preprocess / window grid / PrithviSeg forward: pinned against the reference's
  own modules imported in the dev container (``oracle/refstub.py`` recipe) with
  outputs frozen under ``tests/golden/`` by ``oracle/gen_golden.py``.
* Fmask decode, ``crop_array``, each/any masking: pinned against the reference's
  known-answer tests (``tests/data_tests/test_hls_utils.py:145-159``,
  ``tests/model_tests/test_dataloader.py:117-148``,
  ``tests/data_tests/test_create_chips.py:91-139``).
* timm ``Block`` numerics and the stitch averaging: parity unpinned (no golden
  vector exists in the reference).
"""
