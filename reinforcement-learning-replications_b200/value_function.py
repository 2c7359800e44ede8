from .critics import ValueFunction  # noqa: F401  (module path of the reference API: rl_replicas.value_function)
