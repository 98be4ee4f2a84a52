"""Configuration classes -- the reference's API surface, kept attribute-for-attribute.

Values follow the CODE of the reference, not its README (SURVEY.md trap 1):
/root/reference/wakeword_training_script.py:29-58.  They are plain classes of constants so the
reference's own parametrisation hook keeps working (``class AC(AudioConfig): HOP_LENGTH = 100``).
"""


class AudioConfig:            # wakeword_training_script.py:29-37
    SAMPLE_RATE = 16000
    DURATION = 1.0
    N_MELS = 80
    N_FFT = 2048
    HOP_LENGTH = 512
    WIN_LENGTH = 2048
    FMIN = 0
    FMAX = 8000


class ModelConfig:            # wakeword_training_script.py:39-43
    HIDDEN_SIZE = 256
    NUM_LAYERS = 2
    DROPOUT = 0.6
    NUM_CLASSES = 2


class TrainingConfig:         # wakeword_training_script.py:45-50
    BATCH_SIZE = 16
    LEARNING_RATE = 0.0001
    EPOCHS = 10
    VALIDATION_SPLIT = 0.2
    TEST_SPLIT = 0.1


class AugmentationConfig:     # wakeword_training_script.py:52-58
    AUGMENTATION_PROB = 0.8
    NOISE_FACTOR = 0.15
    TIME_SHIFT_MAX = 0.3
    PITCH_SHIFT_MAX = 3
    SPEED_CHANGE_MIN = 0.7
    SPEED_CHANGE_MAX = 1.3
    # north-star stage set (BASELINE.json): SNR grid of the vendored MS-SNSD synthesiser
    # (stock/ms_snsd/MS-SNSD/noisyspeech_synthesizer.cfg:23-25) for the noise-mix stage
    SNR_GRID_DB = (0.0, 10.0, 20.0, 30.0, 40.0)


class ReadmeAudioConfig(AudioConfig):
    """README figures ("80x160", README.md:176): hop 100 gives 161 frames with centre framing."""
    HOP_LENGTH = 100


class ReadmeModelConfig(ModelConfig):
    """README figure HIDDEN_SIZE = 128 (README.md:150)."""
    HIDDEN_SIZE = 128
