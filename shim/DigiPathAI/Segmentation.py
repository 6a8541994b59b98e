"""``DigiPathAI.Segmentation`` served by the B200 implementation.

The reference module (DigiPathAI/Segmentation.py) exports ``getSegmentation`` (:192-356) and ``get_prediction``
(:65-189); the Flask viewer imports the former by this exact path (main_server.py:155) and calls it with keyword
arguments from a worker thread (main_server.py:165-169).  Same names, same signatures (the union of the three the
reference ships), same ``status`` protocol, same return value -- see digipathai_b200/Segmentation.py.
"""
from digipathai_b200.Segmentation import (get_prediction, getSegmentation,  # noqa: F401
                                          load_trained_models)

__all__ = ["getSegmentation", "get_prediction", "load_trained_models"]
