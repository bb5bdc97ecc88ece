"""crowdnav_b200 -- B200-native batched crowd-navigation environment step.

Public surface:
    crowdnav_b200.config   CnConfig / make_config / baseline_config (worlds as data)
    crowdnav_b200.vec_env  CrowdNavVecEnv  (batched, device tensors)
    crowdnav_b200.env      Env             (the reference's single-env duck type)
    crowdnav_b200.sharded  ShardedVecEnv   (one process per GPU, obs all-gather)
"""
__version__ = "0.1.0"
