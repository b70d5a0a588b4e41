"""ORACLE - test infrastructure only (see ``fusion_decoder.py``).  Never imported by ``transcar_b200``."""
