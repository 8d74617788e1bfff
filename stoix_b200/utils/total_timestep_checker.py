"""Shape derivation for Anakin systems -- same semantics and derived fields as
stoix/utils/total_timestep_checker.py:9-131 (`arch.num_envs`, `arch.num_updates`,
`arch.num_updates_per_eval`), without the colour printing."""
from __future__ import annotations


def check_total_timesteps(config, quiet: bool = False):
    assert config.arch.architecture_name == "anakin", "only the Anakin architecture is built so far"
    n_devices = int(config.num_devices)
    ubs = int(config.arch.update_batch_size)
    divisor = n_devices * ubs
    if int(config.arch.total_num_envs) % divisor != 0:  # :46-54
        raise AssertionError(
            f"The total number of environments ({config.arch.total_num_envs}) must be divisible by "
            f"(num_devices * update_batch_size) = {divisor}!"
        )
    config.arch.num_envs = int(config.arch.total_num_envs) // divisor  # :57-61
    T = int(config.system.rollout_length)
    if config.arch.total_timesteps is None:  # :72-85
        config.arch.total_timesteps = n_devices * int(config.arch.num_updates) * T * ubs * config.arch.num_envs
    else:  # :88-96 -- successive floor divisions, in this order
        config.arch.total_timesteps = int(float(config.arch.total_timesteps))
        config.arch.num_updates = config.arch.total_timesteps // T // ubs // config.arch.num_envs // n_devices
    config.arch.num_updates_per_eval = int(config.arch.num_updates) // int(config.arch.num_evaluation)  # :107
    steps_per_rollout = n_devices * config.arch.num_updates_per_eval * T * ubs * config.arch.num_envs
    actual = steps_per_rollout * int(config.arch.num_evaluation)
    if not quiet and actual != config.arch.total_timesteps:
        print(f"[stoix_b200] total_timesteps={config.arch.total_timesteps:,} is not a multiple of the "
              f"per-evaluation work; {actual:,} steps will be run.")
    return config
