"""Paths used by training_routines (reference: config_template.py:1-6)."""
import os

data_base_path = os.environ.get("RPGP_DATA", os.path.join(os.path.dirname(os.path.abspath(__file__)), "data"))
model_base_path = os.environ.get("RPGP_MODELS", os.path.join(os.path.dirname(os.path.abspath(__file__)), "saved"))
