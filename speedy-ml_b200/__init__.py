"""speedy-ml_b200: B200-native (sm_100a) engine for the SPEEDY-ML local-reservoir hot path.

Import with importlib.import_module("speedy-ml_b200") (the directory name is fixed by the
project layout and is not a Python identifier), or through the alias module speedyml_b200.
"""
