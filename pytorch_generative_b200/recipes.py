"""`reproduce()` of the four autoregressive-image recipes — same signature, hyper-parameters, optimizer, scheduler and loss
as reference models/autoregressive/{pixel_cnn.py:113-176, gated_pixel_cnn.py:193-250, pixel_snail.py:190-262,
image_gpt.py:112-176}, on the B200 path: the model classes of this package, the fused recipe loss, `FusedAdam` and this
package's `Trainer`.  Each model module re-exports its recipe as `reproduce`, like the reference's `train.py` expects.
"""

import torch

from . import losses, optim, trainer


def recipe_loss(x, _, preds):
    """loss_fn(x, _, preds) of every recipe: BCEWithLogits summed over the image, averaged over the batch."""
    return losses.bce_with_logits_sum_mean(preds, x)


def _run(model, lr, lr_gamma, n_epochs, batch_size, log_dir, n_gpus, device_id, debug_loader):
    if n_gpus < 1:
        raise RuntimeError("the B200 path trains on CUDA devices only (n_gpus >= 1); there is no CPU fallback")
    train_loader, test_loader = debug_loader, debug_loader
    if train_loader is None:
        from . import datasets

        device = torch.device("cuda", device_id or 0)
        train_loader, test_loader = datasets.get_mnist_loaders(batch_size, dynamically_binarize=True, device=device)
    optimizer = optim.FusedAdam(model.parameters(), lr=lr)
    scheduler = torch.optim.lr_scheduler.MultiplicativeLR(optimizer, lr_lambda=lambda _: lr_gamma)
    model_trainer = trainer.Trainer(model=model, loss_fn=recipe_loss, optimizer=optimizer, train_loader=train_loader,
                                    eval_loader=test_loader, lr_scheduler=scheduler, log_dir=log_dir, n_gpus=n_gpus,
                                    device_id=device_id)
    model_trainer.interleaved_train_and_eval(n_epochs)
    return model_trainer


def reproduce_pixel_cnn(n_epochs=457, batch_size=256, log_dir="/tmp/run", n_gpus=1, device_id=0, debug_loader=None):
    from . import models

    model = models.PixelCNN(in_channels=1, out_channels=1, n_residual=15, residual_channels=16, head_channels=32)
    return _run(model, 1e-3, 0.999977, n_epochs, batch_size, log_dir, n_gpus, device_id, debug_loader)


def reproduce_gated_pixel_cnn(n_epochs=457, batch_size=128, log_dir="/tmp/run", n_gpus=1, device_id=0, debug_loader=None):
    from . import models

    model = models.GatedPixelCNN(in_channels=1, out_channels=1, n_gated=10, gated_channels=128, head_channels=32)
    return _run(model, 1e-3, 0.9999, n_epochs, batch_size, log_dir, n_gpus, device_id, debug_loader)


def reproduce_pixel_snail(n_epochs=457, batch_size=128, log_dir="/tmp/run", n_gpus=1, device_id=0, debug_loader=None):
    from . import models

    model = models.PixelSNAIL(in_channels=1, out_channels=1, n_channels=64, n_pixel_snail_blocks=8, n_residual_blocks=2,
                              attention_value_channels=32, attention_key_channels=4)
    return _run(model, 1e-3, 0.999977, n_epochs, batch_size, log_dir, n_gpus, device_id, debug_loader)


def reproduce_image_gpt(n_epochs=457, batch_size=64, log_dir="/tmp/run", n_gpus=1, device_id=0, debug_loader=None):
    from . import models

    model = models.ImageGPT(in_channels=1, out_channels=1, in_size=28, n_transformer_blocks=8, n_attention_heads=2,
                            n_embedding_channels=64)
    return _run(model, 5e-3, 0.999977, n_epochs, batch_size, log_dir, n_gpus, device_id, debug_loader)
