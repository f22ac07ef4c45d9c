"""encoder(x, ae_imgcomp, is_training) -> EncoderOutput  (/root/reference/src/encoder_imgcomp.py:4-9)."""


def encoder(x_train, ae_imgcomp, is_training=True, **precision):
    return ae_imgcomp.encode(x_train, is_training=is_training, **precision)
