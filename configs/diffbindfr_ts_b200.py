# DiffBindFR config that routes the hot path through the B200 plugin; everything else is inherited.
# Usage (reference CLI unchanged):  python DiffBindFR/app/predict.py ... -cfg <this file>
# The reference's mmcv-style Config honours `_base_` and `custom_imports` (druglib/utils/config.py:321-328).
_base_ = ['../DiffBindFR/configs/diffbindfr_ts.py']   # adjust to where the reference checkout lives

custom_imports = dict(imports=['diffbindfr_b200.plugin'], allow_failed_imports=False)

model = dict(
    type='DiffBindFRB200',
    diffusion_model=dict(
        type='TensorProductModelB200',
        conv_kernel=6,   # 6: fused tcgen05 FP16x3 on CTA pairs (fp32-grade, default); 5: same on single CTAs (bit-identical); 8: fp16 + e4m3 cross terms on CTA pairs (opt-in, ~5e-5); 4: fused 3xTF32; 0: fp32 SIMT
    ),
)
