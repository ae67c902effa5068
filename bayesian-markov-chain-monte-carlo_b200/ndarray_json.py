"""JSON ndarray codec, wire-compatible with the reference's json_save_load.py.

Format (json_save_load.py:37-38): every ndarray becomes
``{"__ndarray__": true, "data": <tolist()>, "shape": [...]}``; decoding is
``np.array(data).reshape(shape)`` (:128-130).  Like the reference, errors are
printed and swallowed and ``load_object`` returns None on failure (:84-88,
:177-181).  Host-side only; used here for the sample / checkpoint output that
``north_star`` asks for.
"""
import json

import numpy as np


def numpy_array_encoder(obj):
    if isinstance(obj, np.ndarray):
        return {"__ndarray__": True, "data": obj.tolist(), "shape": obj.shape}
    raise TypeError(f"Object of type '{type(obj).__name__}' is not JSON serializable")


def numpy_array_decoder(dct):
    if dct.get("__ndarray__"):
        return np.array(dct["data"]).reshape(dct["shape"])
    return dct


def save_object(obj, filename):
    try:
        with open(filename, "w") as f:
            json.dump(obj, f, default=numpy_array_encoder)
    except Exception as ex:                      # the reference swallows and prints
        print(f"Error during JSON serialization: {ex}")


def load_object(filename):
    try:
        with open(filename, "r") as f:
            return json.load(f, object_hook=numpy_array_decoder)
    except Exception as ex:
        print(f"Error during JSON deserialization: {ex}")
