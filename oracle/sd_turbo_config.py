"""Architecture constants of ``stabilityai/sd-turbo`` (SD-2.1 topology).

TEST INFRASTRUCTURE (oracle).  Everything here is recalled from the public HF
configs / diffusers 0.29 source -- it cannot be
verified offline.  Values corroborated by the reference itself are marked [ref].
Kept in ONE place so they can be corrected the day real ``config.json`` files appear.
"""

UNET = dict(
    in_channels=4,                               # [ref] base_model.py:123
    out_channels=4,
    block_out_channels=(320, 640, 1280, 1280),   # [ref] base_model.py:39
    layers_per_block=2,                          # [ref] 12 skip tensors
    down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),   # [ref] base_model.py:126-144
    up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
    cross_attention_dim=1024,                    # [ref] sd_null_emb.pt is [1,77,1024]
    num_attention_heads=(5, 10, 20, 20),         # config key "attention_head_dim"; head_dim 64
    norm_num_groups=32,
    norm_eps=1e-5,
    transformer_norm_eps=1e-6,                   # Transformer2DModel GroupNorm
    time_embed_in=320,                           # [ref] base_model.py:104-106
    time_embed_dim=1280,
    flip_sin_to_cos=True,
    freq_shift=0,
    upcast_attention=False,
)

VAE = dict(
    in_channels=3,
    out_channels=3,
    latent_channels=4,                           # [ref]
    block_out_channels=(128, 256, 512, 512),     # [ref] autoencoder.py:52-54,94-96
    layers_per_block=2,
    norm_num_groups=32,
    resnet_eps=1e-6,
    scaling_factor=0.18215,
)

SCHEDULER = dict(
    num_train_timesteps=1000,
    beta_start=0.00085,
    beta_end=0.012,
    beta_schedule="scaled_linear",
    prediction_type="epsilon",
    timestep_spacing="trailing",                 # [ref] unifie.py:65-68 {249,499,749,999}
    steps_offset=1,
    set_alpha_to_one=False,
    clip_sample=False,
)

# Reference-owned constants
CONTROLLER = dict(                               # controller.py:29-45 (stablesr_config)
    in_channels=4, model_channels=256, out_channels=256, num_res_blocks=2,
    channel_mult=(1, 1, 2, 2), num_heads=4,
    down_block_types=("AttnDownBlock2D",) * 3 + ("DownBlock2D",),
)
SC_CHANS = [320] * 4 + [640] * 3 + [1280] * 5    # base_model.py:39
SC_COND = 256                                    # base_model.py:30
CFRM_STACKS = ((128, 1), (256, 1), (512, 9))     # autoencoder.py:94-96 (width, #NAFBlocks before AdaNAFV2)
TFA_SPECS = ((512, 512, False), (512, 256, False), (512, 128, True))   # autoencoder.py:122-126
