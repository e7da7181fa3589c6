"""Host logic of the loss wrappers (abcnet_b200/loss.py): which head lists take the direct bf16-P8 gradient path, the operand
geometry it shares with TrainEngine.backward, and the refusal to run without a CUDA device. CPU only."""
import pytest
import torch

from abcnet_b200 import loss as L
from oracle import unet_ref


def _logits(heads, B=1, H=4, W=4):
    return [torch.zeros(B, h, H, W) for h in heads]


def test_head_gradient_operand_geometry():
    # K of the data-gradient GEMM: multiples of 16 channels up to 64, multiples of 64 above (train.py backward uses the same rule)
    assert [L.head_grad_planes(h) for h in unet_ref.V2_HEADS] == [2, 2, 2, 2, 2, 48, 8, 8]
    assert L.head_grad_planes(16) == 2 and L.head_grad_planes(17) == 4 and L.head_grad_planes(64) == 8 and L.head_grad_planes(65) == 16
    for h in range(1, 400):
        assert L.head_grad_planes(h) * 8 >= h


def test_direct_p8_path_only_for_the_v2_head_list():
    assert L.p8_loss_supported(_logits(unet_ref.V2_HEADS))
    assert L.p8_loss_supported(_logits([1, 14, 3, 2, 1, 6 * 32, 32, 32]))           # any n_omega % 4 == 0
    assert not L.p8_loss_supported(_logits([1, 14, 3, 2, 1, 6 * 30, 30, 30]))       # n_omega % 4 != 0
    assert not L.p8_loss_supported(_logits([1, 10, 3, 2, 1, 360, 60, 60]))          # other class counts
    assert not L.p8_loss_supported(_logits([1, 14, 3, 2, 1, 5 * 60, 60, 60]))       # other number of bond types


def test_loss_wrappers_refuse_cpu_tensors():
    if torch.cuda.is_available():
        pytest.skip("checks the no-device behaviour")
    logits = _logits(unet_ref.V2_HEADS)
    targets = [torch.zeros_like(z) for z in logits]
    s = torch.zeros(10)
    with pytest.raises(RuntimeError):
        L.loss_forward_backward(s, None, targets, logits)
    dz = [torch.zeros(1, L.head_grad_planes(z.shape[1]), 4, 4, 8, dtype=torch.bfloat16) for z in logits]
    db = [torch.zeros(z.shape[1], dtype=torch.float64) for z in logits]
    with pytest.raises(RuntimeError):
        L.loss_forward_p8(s, None, targets, logits, dz, db)
