from .tp_scatter import (  # noqa: F401
    B200TensorProductScatter,
    TensorProductScatterInterface,
    enable_B200TensorProductScatter,
)
