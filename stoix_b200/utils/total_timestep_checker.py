"""Shape derivation -- same semantics and derived fields as stoix/utils/total_timestep_checker.py: Anakin :9-131
(`arch.num_envs`, `arch.num_updates`, `arch.num_updates_per_eval`), Sebulba :134-287 (`arch.actor.num_envs_per_actor`,
`arch.learner_parallel_env_consumption`, `arch.local_batch_size`, `arch.global_batch_size`, `arch.num_updates`,
`arch.num_updates_per_eval`), without the colour printing."""
from __future__ import annotations


def check_total_timesteps_sebulba(config, quiet: bool = False):
    """total_timestep_checker.py:134-287.  Needs config.num_actor_devices / num_learner_devices / arch.world_size."""
    arch = config.arch
    total_actors = int(config.num_actor_devices) * int(arch.actor.actor_per_device)
    if int(arch.total_num_envs) % total_actors != 0:  # :176-187
        raise AssertionError(f"The total number of environments ({arch.total_num_envs}) must be divisible by "
                             f"(num_actor_devices * actor_per_device) = {total_actors}!")
    per_device = int(arch.total_num_envs) // int(config.num_actor_devices)  # :190-197
    arch.actor.num_envs_per_actor = int(per_device // int(arch.actor.actor_per_device))
    arch.learner_parallel_env_consumption = (arch.actor.num_envs_per_actor * int(arch.actor.actor_per_device)
                                             * int(config.num_actor_devices))  # :200-213
    arch.local_batch_size = int(int(config.system.rollout_length) * arch.learner_parallel_env_consumption)
    arch.global_batch_size = arch.local_batch_size * int(arch.world_size)
    if arch.total_timesteps is None:  # :216-243
        arch.total_timesteps = int(arch.num_updates) * arch.global_batch_size
    else:
        arch.total_timesteps = int(float(arch.total_timesteps))
        arch.num_updates = int(arch.total_timesteps // arch.global_batch_size)
    num_evaluation = max(int(arch.num_evaluation), 1)  # :246-263
    arch.num_updates_per_eval = int(int(arch.num_updates) // num_evaluation)
    actual = arch.global_batch_size * arch.num_updates_per_eval * num_evaluation
    if not quiet and actual != arch.total_timesteps:
        print(f"[stoix_b200] Timestep discrepancy: expected {arch.total_timesteps:,}, actual {actual:,}.")
    if int(arch.num_updates) <= num_evaluation:  # :281-289
        raise AssertionError(f"Number of updates ({arch.num_updates}) must be greater than number of evaluations ({num_evaluation}).")
    if arch.learner_parallel_env_consumption % int(config.num_learner_devices) != 0:
        raise AssertionError(f"Learner parallel env consumption ({arch.learner_parallel_env_consumption}) must be divisible by "
                             f"number of learner devices ({config.num_learner_devices}).")
    return config


def check_total_timesteps(config, quiet: bool = False):
    if config.arch.architecture_name == "sebulba":
        return check_total_timesteps_sebulba(config, quiet)
    assert config.arch.architecture_name == "anakin", f"unknown architecture '{config.arch.architecture_name}'"
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
