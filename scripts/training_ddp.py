#!/usr/bin/env python
"""DDPM training, single box, one process per GPU -- the B200 counterpart of the reference's ``scripts/training_ddp.py``.

Same command line (``json_config  -d device  -e epochs  -p checkpoint  -l learning_rate``) and the same flow
(config -> schedule -> dataset -> DataLoader / DistributedSampler -> DDPM -> ``configure_optimizers(world)`` -> loop of
``zero_grad / training_step / backward / step``, reference lines 120-206); what changes is the backend block:
``ipex.optimize`` + DistributedDataParallel + oneCCL  ->  ``rho_diffusion.b200.optimize`` (flat buffers, fused
AdamW(+EMA), ONE NCCL all-reduce of the gradient buffer per step).  Launch:

    python scripts/training_ddp.py examples/config_spherical_harmonics.json -d cuda -e 10
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/training_ddp.py examples/config_spherical_harmonics.json -d cuda

``--max-steps`` (extension) stops early; ``--dtype fp32`` selects the exact CUDA-core path.  When ``model.kwargs.cond_fn`` is set
(``examples/config_deep_galaxy.json``: ``"MultiEmbeddings"``) the conditioning module is built from the dataset class's discrete
parameter space exactly like ``scripts/inference.py:122-143`` does, so that config trains as well as samples.  Rank 0 writes
``model.pth`` (reference ``state_dict`` layout) every ``training.save_checkpoint_every_n_epochs`` epochs and at the end, plus
``model_ema.pth`` (the EMA shadow weights in the same layout) when ``training.ema_decay > 0``.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rho-diffusion_b200"))

import torch  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

from rho_diffusion import diffusion  # noqa: E402
from rho_diffusion.config import ExperimentConfig  # noqa: E402
from rho_diffusion.registry import registry  # noqa: E402
from rho_diffusion.utils import ddp_setup, parameter_space_to_embeddings, save_model_checkpoint  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("json_config", type=Path, help="Path to a JSON configuration file; see examples folder.")
    parser.add_argument("-d", "--device", dest="device", type=str, default="cuda", help="Compute device [cuda]")
    parser.add_argument("-e", "--epochs", dest="epochs", type=int, default=10, help="Number of epochs")
    parser.add_argument("-p", dest="model_checkpoint_path", required=False, help="model checkpoint to resume from")
    parser.add_argument("-l", "--learning_rate", dest="learning_rate", type=float, default=0.01, help="(unused, as in the reference)")
    parser.add_argument("--max-steps", type=int, default=0, help="stop after this many optimizer steps (0 = full epochs)")
    parser.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--num-workers", type=int, default=8)
    parser.add_argument("--save", default="model.pth", help="checkpoint file written by rank 0 at the end")
    parser.add_argument("--host-data", action="store_true", help="always go through a host DataLoader, even for datasets that can generate on the device")
    parser.add_argument("--sample-during-training", action="store_true",
                        help="run generate() every training.sample_every_n_epochs epochs (ddpm.py:301-310; commented out in the reference script)")
    parser.add_argument("--cuda-graph", action="store_true",
                        help="capture the whole step (gradient all-reduce buckets included) in a CUDA graph; drops the last partial batch")
    args = parser.parse_args()

    from rho_diffusion import b200 as accel          # raises loudly if the CUDA extension is missing on a GPU box

    config = ExperimentConfig.from_json(args.json_config)
    torch.manual_seed(config.training.seed)          # identical weight initialisation on every rank
    world, rank, local_rank = ddp_setup()
    device = torch.device(f"{args.device}:{local_rank}") if args.device.startswith("cuda") else torch.device(args.device)
    if rank == 0:
        print(config)

    schedule = registry.get("schedules", config.noise_schedule.name)(device=device, **config.noise_schedule.kwargs)
    dset_class = registry.get("datasets", config.dataset.name)
    dset = dset_class(**config.dataset.kwargs)
    sampler = None
    if hasattr(dset, "device_loader") and device.type == "cuda" and not args.host_data:
        # synthetic datasets generate their batches ON THE DEVICE (one kernel launch per batch): no worker processes, no H2D copies
        loader = sampler = dset.device_loader(config.training.batch_size, device, rank=rank, world=world)
    elif world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(dataset=dset, num_replicas=world, rank=rank, seed=config.training.seed)
        loader = DataLoader(dset, batch_size=config.training.batch_size, sampler=sampler, num_workers=args.num_workers, pin_memory=True,
                            drop_last=args.cuda_graph)
    else:
        loader = DataLoader(dset, batch_size=config.training.batch_size, shuffle=True, num_workers=args.num_workers, pin_memory=True,
                            drop_last=args.cuda_graph)

    param_space_embeddings = parameter_space_to_embeddings(config.inference.parameter_space).to(device)
    # conditioning module: `model.kwargs.cond_fn` names a registered layer ("MultiEmbeddings") that is built over the dataset
    # CLASS's discrete parameter space, as scripts/inference.py:122-143 of the reference does; same-shape label embeddings (the
    # spherical-harmonics config) need none
    cond_fn = config.model.kwargs.get("cond_fn")
    parameter_space = getattr(dset_class, "parameter_space", None)
    cond_kwargs = {"parameter_space": parameter_space, "embedding_dim": 128} if (cond_fn and parameter_space is not None) else None
    model = diffusion.DDPM(backbone=config.model.name, backbone_kwargs=config.model.kwargs, schedule=schedule,
                           loss_func=config.training.loss_fn, optimizer=config.optimizer.name, opt_kwargs=config.optimizer.kwargs,
                           cond_fn=cond_fn if cond_kwargs else None, cond_fn_kwargs=cond_kwargs,
                           sample_every_n_epochs=config.training.sample_every_n_epochs,
                           save_checkpoint_every_n_epochs=config.training.save_checkpoint_every_n_epochs,
                           sampling_batch_size=len(param_space_embeddings), labels=param_space_embeddings,
                           sample_parameter_space=config.inference.parameter_space if cond_kwargs else None,
                           seed=config.training.seed)
    if args.model_checkpoint_path is not None:
        model.backbone.load_state_dict(torch.load(args.model_checkpoint_path))
    model = model.to(device)
    # from here on the ranks diverge on purpose: Philox noise is keyed by seed + 7919 * rank (model.global_rank follows the process
    # group), and the CPU generator behind random_timesteps continues from seed + rank
    model.global_rank = rank
    model.reseed_per_rank(config.training.seed)
    optimizer = model.configure_optimizers(world)["optimizer"]            # lr * sqrt(world), as in the reference
    model, optimizer = accel.optimize(model, optimizer=optimizer, dtype=torch.bfloat16 if args.dtype == "bf16" else torch.float32,
                                      ema_decay=config.training.ema_decay)
    model.train()

    def save(epoch=None):
        if rank != 0 or not args.save:
            return
        save_model_checkpoint(model.backbone, args.save)                   # reference state_dict layout (utils.py:166-167)
        if config.training.ema_decay > 0:
            root, ext = os.path.splitext(args.save)
            torch.save({k: v.cpu() for k, v in optimizer.shadow_state_dict(model.backbone).items()}, root + "_ema" + ext)
        print("saved", args.save, "" if epoch is None else f"(epoch {epoch + 1})", flush=True)

    steps, t0, graphed = 0, time.time(), None
    every = int(config.training.save_checkpoint_every_n_epochs or 0)
    for epoch in range(args.epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)                                       # a new shuffle per epoch (the reference never calls it)
        model.current_epoch = epoch
        for step, data in enumerate(loader):
            inputs, labels = data[0].to(device, non_blocking=True), data[1].to(device, non_blocking=True)
            if args.cuda_graph:
                if graphed is None:
                    graphed = accel.GraphedTrainStep(model, optimizer, inputs, labels)
                loss = graphed(inputs, labels)
            else:
                optimizer.zero_grad()
                loss = model.training_step([inputs, labels])
                loss.backward()
                optimizer.step()
            steps += 1
            if rank == 0 and (step % 10 == 0 or steps == args.max_steps):
                print("Epoch: %d/%d, NP: %d, step: %d, loss: %.5f, %.1f samples/s" % (
                    epoch + 1, args.epochs, world, step, loss.item(), steps * len(inputs) * world / (time.time() - t0)), flush=True)
            if args.max_steps and steps >= args.max_steps:
                break
        if args.max_steps and steps >= args.max_steps:
            break
        if every > 0 and epoch > 0 and epoch % every == 0:                 # cadence of ddpm.py:301-317
            save(epoch)
        if args.sample_during_training and epoch > 0 and config.training.sample_every_n_epochs > 0 \
                and epoch % config.training.sample_every_n_epochs == 0:
            model.eval()
            model.generate()                                               # batch-sharded across ranks, no communication
            model.train()
    save()
    if graphed is not None:
        graphed.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
