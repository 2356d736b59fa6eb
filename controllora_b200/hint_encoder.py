"""Hint-encoder execution (ControlLoRA.forward, models.py:810-835) — filled in below."""
def hint_encoder_apply(model, x):
    raise NotImplementedError("hint encoder kernels not wired yet")
