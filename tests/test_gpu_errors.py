"""Error behaviour of the C ABI on a GPU box: every misuse returns a negative BSED_E_* code with a message (never a crash,
never a silent fallback); the Python layer raises BsedError."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_workspace_and_call_order_are_checked():
    from bsed_b200 import _lib, engine
    from bsed_b200._lib import Group, ptr, stream_ptr
    lib = _lib.load()
    plan = engine.Plan(engine.make_cfg(), max_clips=2, device="cuda")
    n = plan.n_params
    flat = torch.zeros(n, device="cuda")
    bn = torch.ones(plan.n_bn, device="cuda")
    x = torch.zeros(2, 1255, 128, device="cuda")
    enc = torch.empty(2, 313, 256, device="cuda")
    grp = (Group * 1)()
    grp[0].params, grp[0].bn_buffers, grp[0].num_batches_tracked = flat.data_ptr(), bn.data_ptr(), None
    grp[0].first_clip, grp[0].n_clips = 0, 2
    # workspace one byte short
    rc = lib.bsed_crnn_forward(plan.p, grp, 1, ptr(x), 2, 0, 0, 0, ptr(enc), ptr(plan.ws), plan.ws_bytes - 1, stream_ptr())
    assert rc == -3 and b"workspace" in lib.bsed_last_error()
    # more clips than the plan was built for
    rc = lib.bsed_crnn_forward(plan.p, grp, 1, ptr(x), 3, 0, 0, 0, ptr(enc), ptr(plan.ws), plan.ws_bytes, stream_ptr())
    assert rc == -1 and b"max_clips" in lib.bsed_last_error()
    # groups that do not tile the batch
    grp[0].n_clips = 1
    rc = lib.bsed_crnn_forward(plan.p, grp, 1, ptr(x), 2, 0, 0, 0, ptr(enc), ptr(plan.ws), plan.ws_bytes, stream_ptr())
    assert rc == -1
    # BSED_F_SAVE without BSED_F_TRAIN
    grp[0].n_clips = 2
    rc = lib.bsed_crnn_forward(plan.p, grp, 1, ptr(x), 2, 2, 0, 0, ptr(enc), ptr(plan.ws), plan.ws_bytes, stream_ptr())
    assert rc == -1
    # backward without a saved forward
    grads = torch.empty(n, device="cuda")
    rc = lib.bsed_crnn_backward(plan.p, 1, ptr(enc), ptr(grads), 0, ptr(plan.ws), plan.ws_bytes, stream_ptr())
    assert rc == -4 and b"saved forward" in lib.bsed_last_error()
    torch.cuda.synchronize()


def test_bad_configurations_are_rejected_at_plan_creation():
    from bsed_b200 import _lib, engine
    for kw, frag in ((dict(nb_filters=(16, 32, 64, 128, 128, 128, 96)), "128 channels"),
                     (dict(nb_filters=(16, 32, 64, 128, 128, 128, 200)), "filters"),
                     (dict(n_RNN_cell=64), "rnn_hidden"),
                     (dict(pooling=((2, 2), (2, 2), (1, 2), (1, 2), (1, 2), (1, 2), (1, 1))), "frequency axis"),
                     (dict(nclass=21), "n_class")):
        with pytest.raises(_lib.BsedError) as e:
            engine.Plan(engine.make_cfg(**kw), max_clips=1, device="cuda", with_workspace=False)
        assert frag in str(e.value), (kw, str(e.value))


def test_python_layer_refuses_cpu_tensors_and_unsupported_kwargs():
    from bsed_b200 import engine
    from bsed_b200.models import CRNN, CRNN_fpn, Predictor
    m = CRNN(**engine.REFERENCE_CRNN_KWARGS).cuda()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 1255, 128))                      # CPU input: no fallback
    with pytest.raises(NotImplementedError):
        CRNN(**{**engine.REFERENCE_CRNN_KWARGS, "activation": "relu"})
    with pytest.raises(NotImplementedError):
        CRNN_fpn(**{**engine.REFERENCE_CRNN_KWARGS, "nb_filters": [16, 32, 64, 128, 128, 128, 64]})
    with pytest.raises(NotImplementedError):
        Predictor(nclass=20, attention=False, n_RNN_cell=128)
    with pytest.raises(ValueError):
        from bsed_b200.models import Clip_Discriminator
        Clip_Discriminator().cuda()(torch.zeros(2, 100, 256, device="cuda"))


def test_frontend_and_decoder_argument_checks():
    from bsed_b200 import _lib, engine
    with pytest.raises(_lib.BsedError):
        engine.melspec(torch.zeros(1, 1000, device="cuda"))   # shorter than the reflect padding
    lib = _lib.load()
    h = _lib.handle(0)
    p = torch.rand(1, 2000, 20, device="cuda")
    ev = torch.zeros(1, 10, 3, dtype=torch.int32, device="cuda")
    n = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = lib.bsed_median_decode(h, _lib.ptr(p), 1, 2000, 20, 0.5, 14, _lib.ptr(ev), 10, _lib.ptr(n), _lib.stream_ptr())
    assert rc == -1                                           # T > 1024
