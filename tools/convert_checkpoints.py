"""Convert the reference's shipped checkpoints into this repo's ``weights/*.npz`` format.

Run in the build container (needs /root/reference); the outputs are committed so that the
GPU box -- which has no /root/reference -- can load real model weights.

    python tools/convert_checkpoints.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from sevenn_b200.checkpoint import convert_reference_checkpoint, save_weights  # noqa: E402

REF = '/root/reference/sevenn/pretrained_potentials'
MODELS = {
    'sevennet_0': f'{REF}/SevenNet_0__11Jul2024/checkpoint_sevennet_0.pth',
    'sevennet_l3i5': f'{REF}/SevenNet_l3i5/checkpoint_l3i5.pth',
}

if __name__ == '__main__':
    out_dir = os.path.join(os.path.dirname(__file__), '..', 'weights')
    os.makedirs(out_dir, exist_ok=True)
    for name, path in MODELS.items():
        meta, arrays = convert_reference_checkpoint(path, name)
        dst = os.path.join(out_dir, f'{name}.npz')
        save_weights(dst, meta, arrays)
        n = sum(a.size for a in arrays.values())
        print(f'{name}: {n} parameters -> {dst} ({os.path.getsize(dst) / 1e6:.2f} MB)')
