"""Mirror of the reference's ``pytorch_models`` package: same class names, constructor arguments,
``forward(data)`` contract and ``state_dict`` keys; the arithmetic runs in the sm_100a kernels."""
