"""B200-native implementation of RFNet's recurrent fusion + decode path."""
from . import models  # noqa: F401
from .model import (AttentionModelCore, FeatArrayFusionNoInputCore, LSTMFusionNoInputCore,  # noqa: F401
                    LSTMSoftAttentionCore, LSTMSoftAttentionNoInputCore,
                    LSTMSoftMultiAttentionFeatArrayNoInputCore, RecurrentFusionModel)
from .models import setup  # noqa: F401
from .options import make_opt  # noqa: F401
