# DiffBindFR config that routes the hot path through the B200 plugin; everything else is inherited.
# Usage (reference CLI unchanged):  python DiffBindFR/app/predict.py ... -cfg <this file>
# The reference's mmcv-style Config honours `_base_` and `custom_imports` (druglib/utils/config.py:321-328).
_base_ = ['../DiffBindFR/configs/diffbindfr_ts.py']   # adjust to where the reference checkout lives

custom_imports = dict(imports=['diffbindfr_b200.plugin'], allow_failed_imports=False)

model = dict(
    type='DiffBindFRB200',
    diffusion_model=dict(
        type='TensorProductModelB200',
        conv_kernel=11,   # 11: fused tcgen05 FP16x3 on CTA pairs + look-ahead gather warpgroup (fp32-grade, default); 6 / 5: the same without look-ahead / on single CTAs (bit-identical); 0: fp32 SIMT
    ),
)
