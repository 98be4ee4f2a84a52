"""wakeword_jupyterlab_b200 -- B200-native drop-in for the hot path of sarpel/wakeword-jupyterlab:
batched 16 kHz clips -> (augment) -> log-mel -> CNN+LSTM score.  See DESIGN.md / INTEGRATION.md."""
from .config import (AudioConfig, AugmentationConfig, ModelConfig, ReadmeAudioConfig, ReadmeModelConfig,
                     TrainingConfig)
from .engine import AugBatch, Engine, get_engine
from .model import WakewordModel
from .predict import predict_wakeword, score_clips, score_stream
from .processor import AudioProcessor
from .trainer import WakewordTrainer
from .checkpoint import (load_checkpoint, load_deployment_package, save_best_checkpoint, save_deployment_package,
                         save_final_checkpoint)
from .dataset import DeviceFeatureLoader, WakewordDataset

__all__ = ["AudioConfig", "ModelConfig", "TrainingConfig", "AugmentationConfig", "ReadmeAudioConfig",
           "ReadmeModelConfig", "AudioProcessor", "WakewordModel", "predict_wakeword", "score_clips",
           "score_stream", "AugBatch", "Engine", "get_engine", "WakewordTrainer", "load_checkpoint", "save_best_checkpoint",
           "save_final_checkpoint", "save_deployment_package", "load_deployment_package", "WakewordDataset",
           "DeviceFeatureLoader"]
