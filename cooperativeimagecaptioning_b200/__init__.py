"""B200-native (sm_100a) joint speaker-listener training path of CooperativeImageCaptioning.

Drop-in for the reference's `models` package on that one path: `setup(opt, model_name, model_type)`
returns the Att2in2 speaker / VSEFC listener whose arithmetic runs in libcoopcap.so.
"""
__version__ = "0.1.0"
