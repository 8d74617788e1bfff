"""Activation lookup (stoix/networks/utils.py:7-25).  The CUDA MLP kernels implement relu, which is
what every PPO network config of the hot path uses (configs/network/mlp.yaml); other names are
recognised so a config error is explicit rather than silent."""

_SUPPORTED = {"relu"}
_KNOWN = {"relu", "tanh", "silu", "elu", "gelu", "sigmoid", "softplus", "swish", "identity", "none",
          "normalise", "softmax", "log_softmax", "log_sigmoid"}


def parse_activation_fn(activation_fn_name: str) -> str:
    if activation_fn_name not in _KNOWN:
        raise KeyError(activation_fn_name)
    if activation_fn_name not in _SUPPORTED:
        raise NotImplementedError(
            f"activation '{activation_fn_name}' is outside the B200 hot path (only relu MLP torsos are built)"
        )
    return activation_fn_name
