"""decoder(z, ae_imgcomp, is_training) -> x_out  (/root/reference/src/decoder_imgcomp.py:3-7)."""


def decoder(z, ae_imgcomp, is_training=True, **precision):
    return ae_imgcomp.decode(z, is_training=is_training, **precision)
