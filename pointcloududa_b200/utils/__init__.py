from .loss import batch_NN_loss, entropy_map  # noqa: F401
