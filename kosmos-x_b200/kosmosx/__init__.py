"""kosmosx — B200-native drop-in for the forward path of kyegomez/Kosmos-X
(same exports as /root/reference/kosmosx/__init__.py:1-4)."""
from kosmosx.model import Decoder, Kosmos, KosmosConfig, KosmosLanguage, KosmosTokenizer
from kosmosx.train import KosmosTrainer

__all__ = ["KosmosTokenizer", "Kosmos", "KosmosLanguage", "Decoder", "KosmosConfig", "KosmosTrainer"]
