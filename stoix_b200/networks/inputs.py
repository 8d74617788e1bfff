"""ArrayInput (stoix/networks/inputs.py:7-12): identity input layer."""


class ArrayInput:
    def __call__(self, embedding):
        return embedding


class EmbeddingActionInput:
    """Observation / embedding and action input: concatenation along the feature axis (stoix/networks/inputs.py:26-33)."""

    def __call__(self, embedding, action):
        import torch

        return torch.cat([embedding, action.to(embedding.dtype)], dim=-1)
