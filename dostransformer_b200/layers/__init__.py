from .transformer import TransformerEncoder

__all__ = ["TransformerEncoder"]
