"""Host-side mirror of the reference's module interface for the hot path
(reference: models/transformers.py, models/encoder.py, models/infonce.py,
models/contrastive.py, models/loss.py, models/fds.py)."""
