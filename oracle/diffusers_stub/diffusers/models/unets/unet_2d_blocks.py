"""get_down_block / get_up_block / UNetMidBlock2D with the argument lists vae.py:104-130,250-285 passes, built from the
oracle's restated blocks (oracle/vq_model_ref.py)."""
from oracle.vq_model_ref import RefDownBlock, RefMidBlock, RefUpBlock


class AutoencoderTinyBlock:     # imported by vae.py, never used by the iVideoGPT path
    pass


def get_down_block(down_block_type, num_layers, in_channels, out_channels, add_downsample, resnet_eps, downsample_padding,
                   resnet_act_fn, resnet_groups, attention_head_dim, temb_channels, **kw):
    assert down_block_type == "DownEncoderBlock2D" and resnet_eps == 1e-6 and downsample_padding == 0
    assert resnet_act_fn == "silu" and temb_channels is None
    return RefDownBlock(in_channels, out_channels, num_layers, add_downsample, resnet_groups)


def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, add_upsample, resnet_eps,
                 resnet_act_fn, resnet_groups, attention_head_dim, temb_channels, resnet_time_scale_shift="default", **kw):
    assert up_block_type == "UpDecoderBlock2D" and resnet_eps == 1e-6 and resnet_act_fn == "silu" and temb_channels is None

    class _Up(RefUpBlock):
        def forward(self, x, temb=None):            # Decoder.forward calls up_block(sample, latent_embeds)
            assert temb is None
            return super().forward(x)
    return _Up(in_channels, out_channels, num_layers, add_upsample, resnet_groups)


class UNetMidBlock2D(RefMidBlock):
    def __init__(self, in_channels, resnet_eps, resnet_act_fn, output_scale_factor, resnet_time_scale_shift,
                 attention_head_dim, resnet_groups, temb_channels, add_attention=True, **kw):
        assert resnet_eps == 1e-6 and resnet_act_fn == "silu" and output_scale_factor == 1 and temb_channels is None
        assert attention_head_dim == in_channels        # one head of dim C
        super().__init__(in_channels, resnet_groups, add_attention)

    def forward(self, x, temb=None):                    # Decoder.forward calls mid_block(sample, latent_embeds)
        assert temb is None
        return super().forward(x)
