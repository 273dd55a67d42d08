"""Shim of diffusers.models.unet_2d_blocks for the three symbols vae.py:22-23 imports."""
from ladi_oracle.vae import DownEncoderBlock2D, UpDecoderBlock2D, VaeMidBlock


def get_down_block(down_block_type, num_layers, in_channels, out_channels, add_downsample, resnet_eps,
                   resnet_act_fn, resnet_groups=32, downsample_padding=0, attn_num_head_channels=None,
                   temb_channels=None, **kw):
    assert down_block_type == "DownEncoderBlock2D" and downsample_padding == 0 and resnet_act_fn == "silu"
    return DownEncoderBlock2D(in_channels, out_channels, num_layers, resnet_groups, resnet_eps, add_downsample)


def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, add_upsample,
                 resnet_eps, resnet_act_fn, resnet_groups=32, attn_num_head_channels=None, temb_channels=None, **kw):
    assert up_block_type == "UpDecoderBlock2D" and resnet_act_fn == "silu"
    return UpDecoderBlock2D(in_channels, out_channels, num_layers, resnet_groups, resnet_eps, add_upsample)


def UNetMidBlock2D(in_channels, resnet_eps, resnet_act_fn, output_scale_factor=1, resnet_time_scale_shift="default",
                   attn_num_head_channels=None, resnet_groups=32, temb_channels=None, **kw):
    assert attn_num_head_channels is None and temb_channels is None
    return VaeMidBlock(in_channels, resnet_groups, resnet_eps)
