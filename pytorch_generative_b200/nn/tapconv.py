"""Wide-channel masked / cropped convolutions as a tap-list GEMM (CausalConv2d with Cin > a few channels,
the GatedPixelCNN 1xN / Nx1 and PixelSNAIL 2x2 convs).  Not built yet in this round: fails loudly."""


def tap_conv2d(x, weight, bias, padding, live_mask=None):
    raise NotImplementedError(
        "tap-list tensor-core convolution (CausalConv2d with wide Cin) is not implemented yet on the B200 path"
    )
