from planerecnet_b200.models.functions.funcs import bias_init_with_prob  # noqa: F401
