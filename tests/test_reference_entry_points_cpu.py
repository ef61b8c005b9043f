"""Drop-in at the reference's dataloader / checkpoint entry points (container only: needs /root/reference; skipped on the GPU
box).  The dataset file written by onet_b200.synth (assemble_dataset = the bookkeeping of prepare_data) is read by the
UNMODIFIED `dataloader/simbg4onet_20230209.py::make_dataloader_snr_range`, whose batches have exactly the shape / dtype /
range `onet_b200.Onet.forward` takes; a checkpoint written by onet_b200.synth.save_checkpoint loads into the UNMODIFIED
reference Onet with `load_state_dict(torch.load(f)['net'])` (Train_Onet_on_simclutter_20250407.py:493) and vice versa."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


def test_reference_dataloader_reads_our_dataset_file(tmp_path):
    from onet_b200 import synth
    import_reference()
    argv, sys.argv = sys.argv, ["x"]
    import dataloader.simbg4onet_20230209 as simbg
    sys.argv = argv
    gen = torch.Generator().manual_seed(0)
    groups = []
    for snr in (0, 1, 2):
        f = torch.rand(10, 1, 400, 400, generator=gen)
        m = torch.rand(10, 400, 400, generator=gen) < 0.01
        groups.append((f, m, snr))
    data = synth.assemble_dataset(groups, (224, 224), "rayleigh")
    assert data["rayleigh_imgs"].shape == (30, 1, 224, 224) and data["rayleigh_labels"].shape == (30, 224, 224)
    # centre crop as torchvision's CenterCrop (the reference's transform, Rayleigh_bg_Gaussian_EOT_generator_20230208.py:296)
    import torchvision.transforms as T
    assert torch.equal(data["rayleigh_imgs"][:10], T.CenterCrop((224, 224))(groups[0][0]))
    torch.save(data, tmp_path / "ours.pt")
    cfg = types.SimpleNamespace(dataset_root=str(tmp_path), data_file_name="ours.pt", preload=True, batch_sz=5,
                                use_augmentation=False, device="cpu")
    np.random.seed(1)
    train_loader, test_loader = simbg.make_dataloader_snr_range(cfg, low_snr=0, high_snr=2)
    assert len(train_loader.dataset) == 27 and len(test_loader.dataset) == 3           # 90 / 10 split (:327-341)
    X, label, snr = next(iter(train_loader))
    assert X.shape == (5, 1, 224, 224) and X.dtype == torch.float32 and 0.0 <= float(X.min()) and float(X.max()) <= 1.0
    assert label.shape == (5, 224, 224) and set(snr.tolist()) <= {0, 1, 2}


def test_checkpoints_interchange_with_the_reference_module(tmp_path):
    import onet_b200
    from onet_b200 import synth
    ref = import_reference()
    torch.manual_seed(4)
    ours = onet_b200.Onet(1, True, True)
    synth.save_checkpoint(ours, 3, str(tmp_path / "ours.pytorch"))
    theirs = ref.Onet(in_chns=1, binit=True, bshare=True)
    theirs.load_state_dict(torch.load(tmp_path / "ours.pytorch", map_location=lambda storage, loc: storage)["net"])   # :493
    for (k1, v1), (k2, v2) in zip(ours.state_dict().items(), theirs.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), (k1, k2)
    torch.save({"net": theirs.state_dict(), "epoch": 7}, tmp_path / "theirs.pytorch")                                  # :264-266
    ours2 = onet_b200.Onet(1, False, True)
    assert synth.load_checkpoint(ours2, str(tmp_path / "theirs.pytorch")) == 7
    for (k1, v1), (_, v2) in zip(ours2.state_dict().items(), theirs.state_dict().items()):
        assert torch.equal(v1, v2), k1
