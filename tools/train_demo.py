"""End-to-end use of the path the way the reference's `unsupervised_training_simclutter` uses it
(Train_Onet_on_simclutter_20250407.py:176-266), entirely on the GPU:

    frames synthesised on the device (onet_b200.synth.prepare_data: Rayleigh clutter + 20 extended targets, peak SNR range)
    -> 90 / 10 split as simbg4onet_20230209.py:327-341 -> OnetTrainer.step (graph replay) for N epochs
    -> evaluate.test_simclutter before / after training -> reference-format checkpoint.

    python tools/train_demo.py [--snr 6 10] [--frames-per-snr 96] [--epochs 6] [--batch 16] [--lr 5e-6]

Prints one JSON line: loss per epoch, the five segmentation metrics before and after, images/s of the training loop.
A demonstration that the kernels train (the loss falls, the unsupervised masks move towards the labels), not a benchmark
and not a reproduction of the paper's 300-epoch runs."""
import argparse
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--snr", type=int, nargs=2, default=[6, 10])
    ap.add_argument("--frames-per-snr", type=int, default=96)
    ap.add_argument("--epochs", type=int, default=6)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--lr", type=float, default=5e-6)          # Train_Onet_on_simclutter_20250407.py:181
    ap.add_argument("--out", default=None, help="write a reference-format checkpoint here")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "tf32", "fp32"])
    args = ap.parse_args()
    import onet_b200
    import onet_b200.evaluate as oev
    from onet_b200 import synth
    from onet_b200.trainer import OnetTrainer

    dev = torch.device("cuda:0")
    torch.manual_seed(1981)
    np.random.seed(1981)
    t0 = time.perf_counter()
    data = synth.prepare_data(img_sz=(224, 224), bg_type="rayleigh", fnums=args.frames_per_snr,
                              snrs=range(args.snr[0], args.snr[1] + 1), seed=1981, device=dev)
    torch.cuda.synchronize()
    t_synth = time.perf_counter() - t0
    imgs, labels, snrs = data["rayleigh_imgs"], data["rayleigh_labels"], torch.tensor(data["psnr"])
    n = imgs.shape[0]
    ids = np.arange(n)
    np.random.shuffle(ids)
    ntrain = int(n * 0.9)
    B = args.batch
    tr_idx, te_idx = ids[:ntrain - ntrain % B], ids[ntrain:]
    test_loader = [(imgs[te_idx[i:i + B]], labels[te_idx[i:i + B]].long(), snrs[te_idx[i:i + B]]) for i in range(0, len(te_idx), B)]
    cfg = types.SimpleNamespace(device=dev)

    net = onet_b200.Onet(1, True, True, mode=args.mode).to(dev)
    trainer = OnetTrainer(net, lr=args.lr, graph=True)
    before = oev.test_simclutter("before", cfg, net, test_loader, verbose=0)
    train_x = imgs[tr_idx].pin_memory()
    loss_epochs, steps = [], 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for epoch in range(args.epochs):
        perm = torch.randperm(train_x.shape[0])
        losses = []
        for i in range(0, len(perm), B):
            losses.append(trainer.step(train_x[perm[i:i + B]]).item())       # loss.item() every step like the reference loop
            steps += 1
        loss_epochs.append(float(np.mean(losses)))
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    after = oev.test_simclutter("after", cfg, net, test_loader, verbose=0)
    if args.out:
        synth.save_checkpoint(net, args.epochs - 1, args.out)
    names = ("acc", "miou", "dr", "far", "tiou")
    print(json.dumps(dict(mode=args.mode, frames=n, train_frames=len(tr_idx), test_frames=len(te_idx), snr=args.snr, batch=B, lr=args.lr,
                          epochs=args.epochs, steps=steps, synth_seconds=round(t_synth, 3), train_seconds=round(t_train, 3),
                          train_images_per_s=round(steps * B / t_train, 1), loss_per_epoch=[round(v, 5) for v in loss_epochs],
                          before=dict(zip(names, (round(v, 4) for v in before))),
                          after=dict(zip(names, (round(v, 4) for v in after))))), flush=True)


if __name__ == "__main__":
    main()
