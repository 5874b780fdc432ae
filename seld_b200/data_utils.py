"""Helpers the extractor shares with its callers (reference data_utils.py:6-17)."""
import os

import numpy as np


def create_folder(folder_name):
    if not os.path.isdir(folder_name):
        print(f'{folder_name} folder does not exist, creating it.')
        os.makedirs(folder_name, exist_ok=True)


def degree_to_radian(degree):
    return degree * np.pi / 180


def radian_to_degree(radian):
    return radian / np.pi * 180
