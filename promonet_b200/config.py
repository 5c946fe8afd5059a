"""Constants of the hot path (values of promonet/config/defaults.py + static.py
under config/promonet.py; one configuration, fixed at import like the reference)"""
RANDOM_SEED = 1234            # defaults.py:120
SAMPLE_RATE = 22050           # :49
HOPSIZE = 256                 # :31
NUM_FFT = 1024                # :43
WINDOW_SIZE = 1024            # :52
NUM_MELS = 80                 # :40
FMIN = 50.                    # :27
FMAX = 550.                   # :28
MIN_DB = -100.                # :37
REF_DB = 20.                  # :46
LOUDNESS_BANDS = 8            # :90
PITCH_BINS = 256              # :96
PITCH_EMBEDDING_SIZE = 64     # :99
PPG_CHANNELS = 40             # :102
SPARSE_PPG_METHOD = 'percentile'  # :113
SPARSE_PPG_THRESHOLD = 0.85   # :117
LRELU_SLOPE = 0.1             # :216
SPEAKER_CHANNELS = 256        # :265
NUM_SPEAKERS = 109            # static.py:58-59 (vctk)
GLOBAL_CHANNELS = 258         # static.py:40-43
NUM_FEATURES = 113            # static.py:47-52
HIFIGAN_RESBLOCK_KERNEL_SIZES = (3, 7, 11)   # :250
HIFIGAN_RESBLOCK_DILATION_SIZES = (1, 3, 5)  # :253
HIFIGAN_UPSAMPLE_INITIAL_SIZE = 512          # :256
HIFIGAN_UPSAMPLE_KERNEL_SIZES = (16, 16, 4, 4)  # :259
HIFIGAN_UPSAMPLE_RATES = (8, 8, 2, 2)        # :262
