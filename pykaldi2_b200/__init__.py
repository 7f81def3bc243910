"""pykaldi2_b200: the sequence-discriminative training hot path of PyKaldi2 on B200 (sm_100a)."""
import os

# The step runs kernels of up to ten CUDA streams side by side (two half-batches, the weight-gradient GEMMs, the
# numerator forward-backward, the denominator's single-CTA kernels, the input prefetch).  With the default of 8
# hardware work queues, streams alias onto the same queue and a long kernel of one stream (a persistent denominator
# CTA) holds back the unrelated kernels queued behind it (measured: profiles/exp_two_halves_r2_v1.jsonl, the
# staggered half-batches ran fully serialised).  Must be set before the CUDA context is created.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
