"""b200med -- B200-native hot path of GonzaloPlaaza/Multimodal-Error-Detection.

Package layout mirrors the reference's ``MED`` package for the path it replaces:

    multimodal_error_detection_b200.dataset.dataset_utils      <- MED/dataset/dataset_utils.py
    multimodal_error_detection_b200.dataset.CustomWindowDataset<- MED/dataset/CustomWindowDataset.py
    multimodal_error_detection_b200.dataset.CustomFrameDataset <- MED/dataset/CustomFrameDataset.py
    multimodal_error_detection_b200.modeling.models            <- MED/modeling/models.py
    multimodal_error_detection_b200.modeling.models_TCN        <- MED/modeling/models_TCN.py (TeCNo)
    multimodal_error_detection_b200.modeling.modeling_utils    <- MED/modeling/modeling_utils.py

``csrc/`` holds the CUDA kernels and the C ABI (include/b200med.h); ``ops`` is the tensor-level
binding.  Nothing here imports ``oracle/`` and nothing falls back to the CPU.
"""
__version__ = "0.1.0"
