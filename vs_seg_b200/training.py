"""Native training path (forward with batch statistics + backward kernels).  Not built yet: the
CUDA train-mode forward raises instead of silently running eager torch."""


def unet_train_forward(model, x):
    raise NotImplementedError(
        "train-mode forward of UNet2d5_spvPA on CUDA needs the native backward kernels, which are not "
        "built yet; there is deliberately no eager-torch CUDA fallback (use model.eval() for inference)")
