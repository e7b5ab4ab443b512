"""tise_toolbox_b200 -- B200 (sm_100a) implementation of TISE's image-realism hot path.

Reference-facing modules (same names and call signatures as the reference's scripts):
  fid_score, inception, img_data                      image_realism/FID, object_fidelity/O-FID
  object_centric_inception_score                      object_fidelity/O-IS
  inception_score_star_coco, inception_score_star_bird image_realism/IS/{coco,bird}
They call hand-written CUDA kernels in libtise_b200.so (csrc/, C ABI in include/tise_b200.h)
through ``engine.Engine``.  There is no CPU fallback and no dependency on ``oracle/``.
"""
__version__ = "0.1.0"
