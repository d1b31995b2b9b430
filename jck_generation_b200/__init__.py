"""jck_generation_b200: B200-native DCGAN / CGAN train step behind the reference's Python API.

Layout mirrors hy-vision-learning/jck-generation (model/, train/, preprocess/, metrics.py, logger/,
utils.py, enums.py, main.py); underneath, `ops` calls the sm_100a kernels of libjck_b200.so through
the C ABI in include/jck_b200.h."""
__version__ = "0.1.0"
