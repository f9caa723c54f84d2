"""Alias package: lets reference code that does ``import dptr.gs as gs`` (src/trainer_fragGS.py:29,
src/pointrix/renderer/dptr*.py:3) bind to the B200-native operators without modification."""
