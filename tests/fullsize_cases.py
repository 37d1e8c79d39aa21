"""Full-size parity cases (BASELINE.json shapes and step counts), shared by the
fixture generator (tools/make_golden_full.py, runs the unmodified reference on the CPU) and the `-m gpu`
parity tests (tests/test_fullsize_parity_gpu.py).  No reference import here."""
from zipvoice_b200.config import ZipVoiceConfig, tiny_config

DIALOG = dict(vocab_size=362)

# name -> dict(cfg, weights, ukw, skw, vel_steps, vel_stride, fm (record one decoder forward))
CASES = {
    # C1 == one utterance of C3: 16 CFG steps, steps 0-10 have t <= 0.5, 11-15 t > 0.5
    "full_c1_zipvoice_16step": dict(
        cfg=ZipVoiceConfig("zipvoice"), weights="synth",
        ukw=dict(batch=1, prompt_frames=281, target_frames=937, prompt_tokens=45, tokens=150),
        skw=dict(num_step=16, guidance_scale=1.0, t_shift=0.5),
        vel_steps=[0, 10, 11, 15], vel_stride=1, fm=True),
    # ragged C3 rows: padded frames, the SimpleDownsample edge (zipformer.py:899-901) at full length
    "full_c3_ragged3_16step": dict(
        cfg=ZipVoiceConfig("zipvoice"), weights="synth",
        ukw=dict(batch=3, prompt_frames=281, target_frames=938, prompt_tokens=45, tokens=150, ragged=True),
        skw=dict(num_step=16, guidance_scale=1.0, t_shift=0.5),
        vel_steps=[0, 10, 11, 15], vel_stride=4, fm=False),
    "full_c2_distill_4step": dict(
        cfg=ZipVoiceConfig("zipvoice_distill"), weights="synth",
        ukw=dict(batch=2, prompt_frames=281, target_frames=938, prompt_tokens=45, tokens=150, ragged=True),
        skw=dict(num_step=4, guidance_scale=3.0, t_shift=0.5),
        vel_steps=[0, 1, 2, 3], vel_stride=2, fm=True),
    # 4 steps with t_shift 0.5: t = 0, .143, .333, .6 -> both CFG branches
    "full_c5_stereo_4step": dict(
        cfg=ZipVoiceConfig("zipvoice_dialog_stereo", **DIALOG), weights="synth",
        ukw=dict(batch=1, prompt_frames=469, target_frames=1875, prompt_tokens=60, tokens=300),
        skw=dict(num_step=4, guidance_scale=1.5, t_shift=0.5),
        vel_steps=[0, 2, 3], vel_stride=4, fm=True),
    # 60 s dialog: t = 0.4 (<= 0.5) and 0.6 (> 0.5)
    "full_c4_dialog_2step": dict(
        cfg=ZipVoiceConfig("zipvoice_dialog", **DIALOG), weights="synth",
        ukw=dict(batch=1, prompt_frames=938, target_frames=5625, prompt_tokens=150, tokens=900),
        skw=dict(num_step=2, guidance_scale=1.5, t_start=0.4, t_end=0.8, t_shift=1.0),
        vel_steps=[0, 1], vel_stride=4, fm=True),
    # the reference's own initialisation: weights not chosen by this repository; the test rebuilds
    # the state_dict from the staged reference package (baseline/_ref) and checks its sha256 against the fixture
    "full_refinit_tiny": dict(
        cfg=tiny_config("zipvoice"), weights="refinit",
        ukw=dict(batch=3, prompt_frames=60, target_frames=200, prompt_tokens=12, tokens=40, ragged=True),
        skw=dict(num_step=8, guidance_scale=1.0, t_shift=0.5),
        vel_steps=[0, 5, 6, 7], vel_stride=1, fm=True),
    "full_refinit_base_16step": dict(
        cfg=ZipVoiceConfig("zipvoice"), weights="refinit",
        ukw=dict(batch=1, prompt_frames=281, target_frames=937, prompt_tokens=45, tokens=150),
        skw=dict(num_step=16, guidance_scale=1.0, t_shift=0.5),
        vel_steps=[0, 10, 11, 15], vel_stride=2, fm=True),
}
