"""Op surface with the reference's signatures (OPS = S3/torch_utils/ops); every op runs sm_100a kernels from
libgantrack_b200.so or raises -- there is no `ref` implementation and no CPU path in this package."""
