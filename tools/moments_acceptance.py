"""Density-field moments acceptance run (north star: "final density-field moments within stated tolerance").

Trains the conditional 2-D UNet of examples/config_gaussian_blobs.json on device-generated Gaussian-blob fields (bf16 tensor-core
path, whole step as one CUDA graph, EMA), then samples with the full T-step ancestral loop from the EMA weights for every point of
the parameter space (K blobs, width w) TWICE with the same Philox keys -- on the bf16 tensor-core path and on the exact fp32 path --
and compares the first two moments of the generated fields per label:

    |mean_bf16 - mean_fp32| <= TOL_MEAN       |std_bf16 - std_fp32| <= TOL_STD       (fields live in [-1, 1])

i.e. the production path reproduces the density-field moments of the reference's arithmetic (whose trajectories the fp32 path
follows to 1e-4 in the moments, tests/test_gpu_model.py).  The dataset's own moments are printed beside them, and the conditioning
must be learned (generated and data mean densities correlate > 0.9 across the labels).  NOTE the generated fields are greyer than
the data (means pulled toward 0 for sparse labels): `DDPM.reverse_process` clamps x_t to [-1, 1] after EVERY step (ddpm.py:218), which
truncates the noisy intermediate states of mostly-background fields -- reference semantics, kept on purpose.
Prints one JSON document (committed under profiles/ with the command line).

    python tools/moments_acceptance.py [--steps 3000] [--per-label 24]
"""
import argparse
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rho-diffusion_b200")]
import torch  # noqa: E402

TOL_MEAN, TOL_STD = 0.02, 0.02


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--per-label", type=int, default=24)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    from rho_diffusion import diffusion
    from rho_diffusion.b200 import GraphedTrainStep, optimize
    from rho_diffusion.config import ExperimentConfig
    from rho_diffusion.registry import registry

    cfg = ExperimentConfig.from_json(os.path.join(ROOT, "examples", "config_gaussian_blobs.json"))
    dev = torch.device("cuda", 0)
    torch.manual_seed(cfg.training.seed)
    dcls = registry.get("datasets", cfg.dataset.name)
    dset = dcls(**cfg.dataset.kwargs)
    schedule = registry.get("schedules", cfg.noise_schedule.name)(device=dev, **cfg.noise_schedule.kwargs)
    model = diffusion.DDPM(backbone=cfg.model.name, backbone_kwargs=cfg.model.kwargs, schedule=schedule, loss_func=cfg.training.loss_fn,
                           timesteps=cfg.noise_schedule.kwargs["num_steps"], optimizer=cfg.optimizer.name, opt_kwargs=cfg.optimizer.kwargs,
                           cond_fn="MultiEmbeddings", cond_fn_kwargs={"parameter_space": dcls.parameter_space, "embedding_dim": 128},
                           seed=cfg.training.seed).to(dev)
    opt = model.configure_optimizers(1)["optimizer"]
    model, fused = optimize(model, optimizer=opt, dtype=torch.bfloat16, ema_decay=0.999)
    model.train()
    loader = dset.device_loader(args.batch, dev)
    x, y = dset.device_batch(args.batch, dev)
    step = GraphedTrainStep(model, fused, x, y)
    t0, losses = time.time(), []
    for i in range(args.steps):
        x, y = dset.device_batch(args.batch, dev)
        loss = step(x, y)
        if i % 250 == 0 or i == args.steps - 1:
            losses.append((i, float(loss)))
    torch.cuda.synchronize()
    train_s = time.time() - t0
    step.close()
    # sample from the EMA weights
    model.backbone.load_state_dict(fused.shadow_state_dict(model.backbone))
    model.eval()
    space = cfg.inference.parameter_space
    combos = list(itertools.product(*space.values()))
    rows = []
    t0 = time.time()
    from rho_diffusion.b200 import set_compute_dtype
    gen_by, gen32_by, data_by = {}, {}, {}
    for dtype, store in ((torch.bfloat16, gen_by), (torch.float32, gen32_by)):
        set_compute_dtype(model, dtype)
        model._samplers = {}
        for ci, combo in enumerate(combos):
            cond = torch.tensor([combo] * args.per_label, dtype=torch.float32, device=dev)
            shape = (args.per_label, 1, *cfg.model.kwargs["data_shape"])
            model._noise_calls = 1000 + 2 * ci                 # the same Philox keys (x_T and every step's z) on both paths
            out = model.reverse_process(torch.zeros(shape, device=dev), conditions=cond)["denoised"]
            store[combo] = (float(out.mean()), float(out.std()))
    sample_s = time.time() - t0
    # data moments per label from the host restatement of the generator
    acc = {c: [] for c in combos}
    i = 0
    while min(len(v) for v in acc.values()) < args.per_label and i < 20000:
        f, lab = dset[i % len(dset)] if i < len(dset) else (None, None)
        if f is None:
            break
        key = (int(lab[0]), float(lab[1]))
        match = [c for c in combos if int(c[0]) == key[0] and abs(float(torch.tensor(c[1], dtype=torch.float32)) - key[1]) < 1e-9]
        if match and len(acc[match[0]]) < 4 * args.per_label:
            acc[match[0]].append(f)
        i += 1
    ok = True
    for c in combos:
        fields = torch.stack(acc[c])
        data_by[c] = (float(fields.mean()), float(fields.std()))
        dm, ds = abs(gen_by[c][0] - gen32_by[c][0]), abs(gen_by[c][1] - gen32_by[c][1])
        good = dm <= TOL_MEAN and ds <= TOL_STD
        ok &= good
        rows.append({"n": c[0], "w": c[1], "bf16_mean": gen_by[c][0], "fp32_mean": gen32_by[c][0], "bf16_std": gen_by[c][1], "fp32_std": gen32_by[c][1],
                     "data_mean": data_by[c][0], "data_std": data_by[c][1], "samples_data": len(acc[c]), "within_tolerance": good})
    gm = torch.tensor([gen_by[c][0] for c in combos]); dmn = torch.tensor([data_by[c][0] for c in combos])
    corr = float(torch.corrcoef(torch.stack([gm, dmn]))[0, 1])
    mono = corr > 0.9
    print(json.dumps({"config": "examples/config_gaussian_blobs.json", "train_steps": args.steps, "batch": args.batch, "train_seconds": train_s,
                      "train_samples_per_s": args.steps * args.batch / train_s, "loss": losses, "sampling_seconds": sample_s,
                      "per_label_samples": args.per_label, "tolerance": {"mean": TOL_MEAN, "std": TOL_STD}, "labels": rows,
                      "all_within_tolerance": ok, "gen_vs_data_mean_correlation": corr, "conditioning_learned": mono}, indent=1))
    sys.exit(0 if (ok and mono) else 1)


if __name__ == "__main__":
    main()
