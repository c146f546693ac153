"""Causal streaming (BASELINE cfg-5): feed a stream in chunks of >= 1 frames.

The network is causal -- each (2,3) encoder conv looks back exactly one frame and the GRUs carry
their hidden state (the reference's state-carry API is ``GroupedGRULayer.forward(input, h0) ->
(out, h)``, model/based_model/cust_conv.py:303-325) -- so a chunked run with this state equals the
batched forward.
"""
from __future__ import annotations

import torch


class StreamState:
    """hist[k]: last input frame of encoder stage k+1, [B,Cin,Fin]; gru: (h1, h2) each [G,B,H]."""

    def __init__(self):
        self.hist = None
        self.gru = None


def step(model, mag_chunk, state: StreamState):
    """mag_chunk [B,Tc,F] (Tc >= 1 new frames of B concurrent utterances) -> mask [B,Tc,F]; updates ``state``."""
    if model.training:
        raise RuntimeError("streaming.step needs model.eval() (batch statistics are undefined on a stream)")
    with torch.no_grad():
        return model.forward_frames(mag_chunk, state=state, want_state=True)
