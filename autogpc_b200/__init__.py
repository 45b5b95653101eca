"""autogpc_b200 -- B200-native candidate-scoring engine behind AutoGPC's GPCKernel / GPCSearch API."""
__version__ = "0.1.0"
