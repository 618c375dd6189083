"""Checkpoint dict of the reference (src/main.py:854-925) and its key-compatibility rules.

    state = {'model': {name, args, kwargs, state_dict}, 'model_p': {...}, 'model_ema': {...}, 'model_p_ema': {...},
             'optimizer': {name, args, kwargs, state_dict}, 'pooling_time_ratio', 'scaler', 'many_hot_encoder',
             'median_window', 'epoch'}

Reference checkpoints of the plain CRNN carry CNN keys with ONE `cnn.` prefix (CNN.state_dict() is overridden,
src/models/CNN.py:71-75) and are loaded through a `cnn.` -> `cnn.cnn.` rename hack (src/main.py:794-798,
src/TestModel.py:48-52).  This package's CRNN uses the 79 canonical keys (`cnn.conv0.weight` ...), so loading accepts
either spelling and saving can emit either.
"""
import torch


def canonical_state_dict(sd):
    """`cnn.cnn.*` (after the reference's rename hack) -> `cnn.*`; other keys unchanged."""
    return {(k.replace("cnn.cnn.", "cnn.", 1) if k.startswith("cnn.cnn.") else k): v for k, v in sd.items()}


def reference_renamed_state_dict(sd):
    """`cnn.*` -> `cnn.cnn.*` (what the reference's own load path builds before load_state_dict)."""
    return {("cnn." + k if k.startswith("cnn.") and not k.startswith("cnn.cnn.") else k): v for k, v in sd.items()}


def _entry(module, kwargs):
    return {"name": type(module).__name__, "args": "", "kwargs": dict(kwargs),
            "state_dict": {k: v.detach().cpu() for k, v in module.state_dict().items()}}


def build_state(model, predictor, crnn_kwargs, predictor_kwargs, optimizer=None, ema_model=None, ema_predictor=None,
                pooling_time_ratio=4, scaler=None, many_hot_encoder=None, median_window=14, epoch=0):
    """The reference's `state` dict (src/main.py:854-925)."""
    state = {"model": _entry(model, crnn_kwargs), "model_p": _entry(predictor, predictor_kwargs),
             "pooling_time_ratio": pooling_time_ratio, "median_window": median_window, "epoch": epoch,
             "scaler": scaler.state_dict() if hasattr(scaler, "state_dict") else scaler,
             "many_hot_encoder": many_hot_encoder.state_dict() if hasattr(many_hot_encoder, "state_dict")
             else many_hot_encoder}
    if ema_model is not None:
        state["model_ema"] = _entry(ema_model, crnn_kwargs)
    if ema_predictor is not None:
        state["model_p_ema"] = _entry(ema_predictor, predictor_kwargs)
    if optimizer is not None:
        state["optimizer"] = {"name": type(optimizer).__name__, "args": "",
                              "kwargs": {k: v for k, v in optimizer.defaults.items()},
                              # FusedAdam emits torch.optim.Adam's layout (step / exp_avg / exp_avg_sq per parameter)
                              "state_dict": optimizer.state_dict()}
    return state


def save_state(state, path):
    torch.save(state, path)


def load_models(state, model, predictor, ema_model=None, ema_predictor=None, optimizer=None):
    """Load a reference-format checkpoint dict into this package's modules (either CNN key spelling) and, when given,
    the optimizer (`optim.load_state_dict(state['optimizer']['state_dict'])`; a FusedAdam takes torch.optim.Adam's
    moments into the trainer's flat buffers, so a resumed run continues with its bias correction and moments)."""
    model.load_state_dict(canonical_state_dict(state["model"]["state_dict"]))
    predictor.load_state_dict(state["model_p"]["state_dict"])
    if ema_model is not None and "model_ema" in state:
        ema_model.load_state_dict(canonical_state_dict(state["model_ema"]["state_dict"]))
    if ema_predictor is not None and "model_p_ema" in state:
        ema_predictor.load_state_dict(state["model_p_ema"]["state_dict"])
    if optimizer is not None and "optimizer" in state:
        optimizer.load_state_dict(state["optimizer"]["state_dict"])
    return state.get("epoch", 0)
