from .models import BigVGAN, VocoderBigVGAN  # noqa: F401
