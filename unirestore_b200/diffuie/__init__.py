"""B200-native mirror of the reference's ``src/modules/diffuie`` operator surface.

Same class names, constructor arguments, attribute paths and state_dict keys as the reference modules
(``DiffUIE``, ``SkipConnectedAutoEncoder``, ``Controller``, ``ControlledUNet``, ``CSCEAdapter``, ``NAFBlock``,
``AdaNAFV2``, ``TaskFeatureAdapter``); every forward launches hand-written sm_100a kernels through the C-ABI.
"""
from .autoencoder import SkipConnectedAutoEncoder
from .base_model import ControlledUNet
from .cfrm import AdaNAFV2
from .controller import Controller, stablesr_config
from .nafnet_arch import NAFBlock
from .scedit import CSCEAdapter
from .taskeditor import TaskFeatureAdapter
from .unifie import DiffUIE

TaskEditorV1c = TaskFeatureAdapter      # pre-release name still imported at autoencoder.py:112 of the reference

__all__ = ["DiffUIE", "SkipConnectedAutoEncoder", "Controller", "ControlledUNet", "CSCEAdapter", "NAFBlock",
           "AdaNAFV2", "TaskFeatureAdapter", "TaskEditorV1c", "stablesr_config"]
