"""The reference's OWN test suite, run against this backend with the same constructor calls, fixtures and assertions
(/root/reference/tests: ``test_activations.py:26-35``, ``models/test_unet.py:28-56``, ``pipeline/test_lightning_ddpm.py:33-90``;
``pipeline/test_schedule.py`` is mirrored in ``test_host_api.py``).  The only change is the device: the reference runs its smoke tests
on the CPU, this backend has no CPU path, so models and inputs live on ``cuda:0`` (tests that compute are marked ``gpu``).

What they pin is API shape, not numerics -- backbone and loss passed as CLASSES, a schedule built positionally, dict batches whose
``"labels"`` key is silently ignored (the code reads ``"label"``, ddpm.py:254: the run is unconditional), ``forward_process`` unpacking into
two tensors, a truthy scalar loss (~ mean(eps^2): the freshly initialised UNet ends in a zeroed conv and predicts exactly 0)."""
import pytest
import torch
from torch import nn

import rho_diffusion  # noqa: F401
from rho_diffusion.diffusion import DDPM, GaussianDiffusionPipeline
from rho_diffusion.models.unet_v2 import UNet
from rho_diffusion.registry import registry

DEV = "cuda:0"


# ---- tests/test_activations.py:26-35 (host modules: torch's own activations and the closed-form SymmetricLog) ----------------------
@pytest.mark.parametrize("classname", list(registry.mapping["activations"]))
@pytest.mark.filterwarnings("ignore:Implicit dimension choice")
def test_activation_functions(classname):
    x = torch.rand(8, 10)
    class_type = registry.get("activations", classname)
    activation = class_type()
    with torch.inference_mode():
        y = activation(x)
    assert y.shape == x.shape and torch.isfinite(y).all()


# ---- tests/models/test_unet.py:28-56 ------------------------------------------------------------------------------------------
def test_unet_in_registry():
    model_class = registry.get("models", "UNetv2")
    assert model_class, "UNetv2 was not found in registry!"
    assert model_class is UNet


def test_unet_init():
    unet = UNet(data_shape=16, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=2)
    assert isinstance(unet, nn.Module) and sum(p.numel() for p in unet.parameters()) > 0


@pytest.mark.gpu
def test_unet_forward():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    torch.manual_seed(21516)
    model = UNet(data_shape=16, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=2).to(DEV)
    x = torch.rand(8, 3, 24, 16, device=DEV)
    timesteps = torch.arange(8, device=DEV)
    with torch.inference_mode():
        pred_y = model(x, timesteps)
    assert isinstance(pred_y, torch.Tensor) and pred_y.shape == x.shape
    assert ~torch.isnan(pred_y).all()
    assert float(pred_y.abs().max()) == 0.0          # zero_module on the last conv (unet_v2.py:650-654): exactly 0, as in the reference


# ---- tests/pipeline/test_lightning_ddpm.py:33-90 -------------------------------------------------------------------------------
KWARGS = {"in_channels": 3, "model_channels": 32, "out_channels": 3, "num_res_blocks": 2, "data_shape": [16, 16]}


def _pipeline(cls):
    torch.manual_seed(21516)
    unet_class = registry.get("models", "UNetv2")
    schedule_class = registry.get("schedules", "LinearSchedule")
    schedule = schedule_class(1000, 1e-4, 0.02)
    return cls(unet_class, dict(KWARGS), schedule, nn.MSELoss)


@pytest.mark.parametrize("cls", [DDPM, GaussianDiffusionPipeline], ids=["ddpm", "gaussian_diffusion"])
def test_pipeline_can_be_instantiated_the_reference_way(cls):
    pipe = _pipeline(cls)
    assert isinstance(pipe.backbone, UNet) and len(pipe.schedule["beta_t"]) == 1000


@pytest.mark.gpu
@pytest.mark.parametrize("cls", [DDPM, GaussianDiffusionPipeline], ids=["ddpm", "gaussian_diffusion"])
def test_pipeline_forward(cls):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    pipe = _pipeline(cls).to(DEV)
    mean, var = pipe.forward_process(torch.rand(1, 3, 16, 16, device=DEV))
    assert torch.isfinite(mean).all()
    assert torch.isfinite(var).all()


@pytest.mark.gpu
@pytest.mark.parametrize("cls", [DDPM, GaussianDiffusionPipeline], ids=["ddpm", "gaussian_diffusion"])
def test_pipeline_train_step(cls):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    pipe = _pipeline(cls).to(DEV)
    test_batch = {"data": torch.rand(1, 3, 16, 16, device=DEV), "labels": torch.LongTensor([1]).to(DEV)}
    loss = pipe.training_step(test_batch)
    assert loss
    assert torch.isfinite(loss) and 0.0 < float(loss.detach()) < 10.0
    loss.backward()                                   # the reference's Trainer would; every parameter that is used gets a finite gradient
    grads = [p.grad for p in pipe.backbone.parameters() if p.grad is not None]
    assert grads and all(torch.isfinite(g).all() for g in grads)
