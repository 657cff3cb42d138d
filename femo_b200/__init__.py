"""femo_b200 -- B200-native finite-element state/adjoint engine behind femo's API."""
__version__ = '0.1.0'
