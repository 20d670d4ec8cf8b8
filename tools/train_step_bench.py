"""cfg4 (BASELINE.json configs[3]): the frequency-aware CycleGAN training step of the reference, batch-sharded over
the GPUs of one box, with the b200wave path patched in or with the stock reference code.  BENCH INFRASTRUCTURE: it
imports the UNMODIFIED reference (model.py, utils.py, pytorch_wavelets, ssim.py) from the git-ignored oracle/_ref
(`make -C oracle ref`); the product never does.

    python tools/train_step_bench.py --impl b200wave [--steps K] [--warmup W]
    python tools/train_step_bench.py --impl reference
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/train_step_bench.py --impl b200wave            # one rank per GPU, its own batch, gradients averaged (NCCL)

The step is train.py:164-269 restated as a function (the reference has it in the script body): the four generator
passes with their Fourier-domain splits (utils.high_pass / low_pass, train.py:173-213), generator losses incl.
`loss_ssim` re-enabled (train.py:234, commented out upstream) and the TV term (train.py:178), AdamW step, the two
discriminator updates through the replay buffers, AdamW step.  Inputs are synthetic (1, 1, 256, 256) crops per rank
(train.py defaults: sizeA * 2 = sizeB = 256, batchSize 1 -- the step is batch-1 by construction: `real_A[0]`).
Networks are cuDNN in both arms.  What changes with --impl b200wave:
  * `from pytorch_wavelets import DWTForward` (model.py:4) and `import ssim` (train.py:30) resolve to b200wave
    (compat.install) -- the discriminators' DWT, SSIM forward + backward;
  * FS_Discriminator*.filter_wavelet -> fused analysis-kernel epilogue, TVLoss -> fused (compat.patch_model);
  * utils.high_pass / low_pass -> batched rfft2 + in-kernel mask (compat.patch_utils) instead of a rows x cols Python
    loop on the host per call.
Multi-GPU: every rank runs the step on its own batch; after each backward the gradients of that optimizer's
parameters are averaged with ONE flat NCCL all-reduce (what DDP does, without its per-forward bookkeeping -- the step
calls each generator three times before one backward, and NetworkA2B.unet is never used).  Timing: CUDA events around
K steps after W warm-up steps, barrier both sides, max over ranks; rank 0 prints one JSON line.
"""
import argparse
import json
import os
import random
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Stub(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def import_reference(impl):
    """Returns (model, utils, ssim) modules: the unmodified reference files, with b200wave patched in for impl='b200wave'."""
    for name in ["skimage", "skimage.metrics", "skimage.io", "skimage.measure", "matplotlib", "matplotlib.pyplot", "cv2",
                 "tqdm", "torchvision", "torchvision.utils", "torchvision.transforms", "torchvision.models", "PIL",
                 "PIL.Image", "visdom", "tkinter"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isfile(os.path.join(ref, "model.py")):
        raise SystemExit("oracle/_ref/model.py is not staged: run `make -C oracle ref` where /root/reference exists")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pywt_standin"))
    if impl == "b200wave":
        import b200wave
        from b200wave import compat
        compat.install(force=True)          # pytorch_wavelets / ssim import names -> b200wave
        sys.path.append(ref)                # model.py, utils.py only (the aliases shadow the staged packages)
        import model
        import utils
        import ssim
        compat.patch_model(model)
        compat.patch_utils(utils)
        assert ssim.SSIM is b200wave.SSIM and model.DWTForward is b200wave.DWTForward
    else:
        sys.path.insert(0, ref)
        import model
        import utils
        import ssim
        assert os.path.abspath(ssim.__file__).startswith(ref) and os.path.abspath(model.__file__).startswith(ref)
    return model, utils, ssim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["b200wave", "reference"], default="b200wave")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=256)
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    saved = os.dup(1)        # the reference's constructors print; keep stdout for the one JSON line
    os.dup2(2, 1)
    model, utils, ssim = import_reference(args.impl)
    torch.manual_seed(0)     # same initial weights on every rank and in both arms
    random.seed(0)
    netG_A2B, netG_B2A = model.NetworkA2B().to(dev), model.NetworkB2A().to(dev)
    netD_A, netD_B = model.FS_DiscriminatorA(1).to(dev), model.FS_DiscriminatorB(1).to(dev)
    for net in (netG_A2B, netG_B2A, netD_A, netD_B):
        net.apply(utils.weights_init_normal)
    crit_gan, crit_cycle, crit_idt = torch.nn.MSELoss(), torch.nn.L1Loss(), torch.nn.L1Loss()
    crit_feature = torch.nn.BCEWithLogitsLoss()
    crit_ssim = ssim.SSIM()
    crit_tv = model.TVLoss().to(dev)
    import itertools
    g_params = list(itertools.chain(netG_A2B.parameters(), netG_B2A.parameters()))
    d_params = list(itertools.chain(netD_A.parameters(), netD_B.parameters()))
    opt_G = torch.optim.AdamW(g_params, lr=1.3e-4, betas=(0.9, 0.999))
    opt_D = torch.optim.AdamW(d_params, lr=1.3e-4, betas=(0.9, 0.999))
    beta1, beta2, beta3, beta4, beta5 = 0.25, 10.0, 2.0, 0.5, 0.5      # train.py:50-54
    target_real = torch.ones(1, device=dev)
    target_fake = torch.zeros(1, device=dev)
    buf_A, buf_B = utils.ReplayBuffer(), utils.ReplayBuffer()

    if args.impl == "b200wave":
        from b200wave.sharding import average_gradients as average_grads   # one flat NCCL all-reduce per optimizer
    else:   # the reference arm imports nothing of the package: the same flat all-reduce, restated
        def average_grads(params):
            if world == 1:
                return
            grads = [p.grad for p in params if p.grad is not None]
            flat = torch._utils._flatten_dense_tensors(grads)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat /= world
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(f)

    def split(img, hi, lo):          # train.py:173-176 and its three repeats
        hf = utils.high_pass(img[0], i=hi).unsqueeze(0).unsqueeze(0)
        hf = (hf + img) / 2.0
        lf = utils.low_pass(img[0], i=lo).unsqueeze(0).unsqueeze(0)
        return hf, lf

    def step(real_A, real_B):
        # (1) forward, train.py:170-213
        hf, lf = split(real_A, 10, 8)
        lf_feature_A, hf_feature_A, fake_B = netG_A2B(lf, hf)
        tv_fake_B = crit_tv(fake_B) * 0.5
        _, _, idt_A = netG_B2A(hf, lf)
        hf_feature_A, lf_feature_A = hf_feature_A.detach(), lf_feature_A.detach()
        hf, lf = split(fake_B, 5, 14)
        hf_feature_recovered_A, _, recovered_A = netG_B2A(hf, lf)
        hf, lf = split(real_B, 5, 14)
        hf_feature_B, lf_feature_B, fake_A = netG_B2A(hf, lf)
        _, _, idt_B = netG_A2B(lf, hf)
        lf_feature_B, hf_feature_B = lf_feature_B.detach(), hf_feature_B.detach()
        hf, lf = split(fake_A, 10, 8)
        _, hf_feature_recovered_B, recovered_B = netG_A2B(lf, hf)
        # (2) generators, train.py:216-240 (+ loss_ssim of :234 and the TV term of :178)
        utils.set_requires_grad([netD_A, netD_B], False)
        opt_G.zero_grad()
        loss_GAN_A2B = crit_gan(netD_B(fake_B), target_real) * beta4
        loss_GAN_B2A = crit_gan(netD_A(fake_A), target_real) * beta5
        loss_cycle_ABA = crit_cycle(recovered_A, real_A) * beta3 + crit_feature(hf_feature_A, hf_feature_recovered_A)
        loss_cycle_BAB = crit_cycle(recovered_B, real_B) * beta3 + beta1 * crit_feature(hf_feature_B, hf_feature_recovered_B)
        loss_idt = crit_idt(real_A, idt_A) * beta2 + crit_idt(real_B, idt_B) * beta2
        loss_ssim = (1 - crit_ssim(recovered_A, real_A)) + (1 - crit_ssim(recovered_B, real_B))
        loss_G = loss_GAN_A2B + loss_GAN_B2A + loss_cycle_ABA + loss_cycle_BAB + loss_idt + loss_ssim + tv_fake_B
        loss_G.backward()
        average_grads(g_params)
        opt_G.step()
        # (3) discriminators, train.py:243-269
        utils.set_requires_grad([netD_A, netD_B], True)
        opt_D.zero_grad()
        loss_D_A = (crit_gan(netD_A(real_A), target_real)
                    + crit_gan(netD_A(buf_A.push_and_pop(fake_A).detach()), target_fake)) * 0.5
        loss_D_A.backward()
        loss_D_B = (crit_gan(netD_B(real_B), target_real)
                    + crit_gan(netD_B(buf_B.push_and_pop(fake_B).detach()), target_fake)) * 0.5
        loss_D_B.backward()
        average_grads(d_params)
        opt_D.step()
        return loss_G.detach(), loss_ssim.detach()

    gen = torch.Generator(device="cpu").manual_seed(100 + rank)      # every rank its own batch
    n = args.size
    batches = [(torch.rand((1, 1, n, n), generator=gen).mul_(2).sub_(1).to(dev),
                torch.rand((1, 1, n, n), generator=gen).mul_(2).sub_(1).to(dev)) for _ in range(4)]
    first = None
    for i in range(args.warmup):
        out = step(*batches[i % len(batches)])
        if first is None:
            first = [float(out[0]), float(out[1])]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(*batches[i % len(batches)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        launches = None
        if args.impl == "b200wave":
            from b200wave import _cabi
            launches = _cabi.kernel_launches()
        os.dup2(saved, 1)
        print(json.dumps({
            "workload": "cfg4: frequency-aware CycleGAN training step (train.py:164-269, loss_ssim + TV enabled), "
                        "1x1x%dx%d A and B crops per GPU, gradients averaged over ranks" % (n, n),
            "impl": args.impl, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "steps_per_s": world * args.steps / (ms * 1e-3),
            "first_step": {"loss_G": first[0], "loss_ssim": first[1]} if first else None,
            "b200wave_kernel_launches_total": launches, "scaling": "weak",
            "torch": torch.__version__}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
