"""HF-independent host models for the model-level goldens, shared by the generator (which fills them with the
REFERENCE's quantizable blocks and runs the reference's quantize()) and by the GPU tests (which fill them with this
repo's blocks).  Only the glue lives here -- Linear / activation / norm modules and the order the blocks are called
in, i.e. what Hugging Face's own *Layer classes do (transformers 4.x MobileBertLayer / BertLayer / LlamaModel); every
quantizable block comes from `blocks`.  Pure torch + transformers: importable on the GPU box.

Why not the HF models themselves: the reference's blocks have transformers-4.x call signatures and cannot be swapped
into the transformers 5.x models of this image."""
import math
from types import SimpleNamespace

import torch
from torch import nn


# ----------------------------------------------------------------------------------------------- MobileBERT
def mobilebert_config(hidden=128, true_hidden=32, heads=4, inter=128, ffn=2):
    from transformers import MobileBertConfig
    return MobileBertConfig(hidden_size=hidden, intra_bottleneck_size=true_hidden, num_attention_heads=heads,
                            intermediate_size=inter, num_feedforward_networks=ffn, hidden_act="relu",
                            normalization_type="no_norm", use_bottleneck=True, use_bottleneck_attention=False,
                            key_query_shared_bottleneck=True, hidden_dropout_prob=0.0,
                            attention_probs_dropout_prob=0.0, embedding_size=32, num_hidden_layers=2)


class MobileBertHostLayer(nn.Module):
    """transformers' MobileBertLayer + MobileBertAttention + FFNLayer call order (bottleneck -> self-attention ->
    self-output -> (num_ffn - 1) x [intermediate, FFNOutput] -> intermediate -> output(+ output bottleneck))."""

    def __init__(self, blocks, cfg):
        super().__init__()
        from transformers.models.mobilebert import modeling_mobilebert as hf
        self.bottleneck = hf.Bottleneck(cfg)
        self.self = blocks.MobileBertSelfAttention(cfg)
        self.self_out = blocks.MobileBertSelfOutput(cfg)
        self.ffn_inter = nn.ModuleList([hf.MobileBertIntermediate(cfg) for _ in range(cfg.num_feedforward_networks - 1)])
        self.ffn_out = nn.ModuleList([blocks.FFNOutput(cfg) for _ in range(cfg.num_feedforward_networks - 1)])
        self.intermediate = hf.MobileBertIntermediate(cfg)
        self.output = blocks.MobileBertOutput(cfg)

    def forward(self, x, mask):
        q, k, v, layer_input = self.bottleneck(x)
        a = self.self(q, k, v, mask)[0]
        a = self.self_out(a, layer_input)
        for inter, out in zip(self.ffn_inter, self.ffn_out):
            a = out(inter(a), a)
        return self.output(self.intermediate(a), a, x)


class MobileBertHost(nn.Module):
    def __init__(self, blocks, cfg, layers=2):
        super().__init__()
        self.layers = nn.ModuleList([MobileBertHostLayer(blocks, cfg) for _ in range(layers)])
        self.qa_outputs = nn.Linear(cfg.hidden_size, 2)

    def forward(self, x, mask):
        for layer in self.layers:
            x = layer(x, mask)
        return self.qa_outputs(x)


# ----------------------------------------------------------------------------------------------- BERT (+ LoRA)
def bert_config(hidden=64, heads=4, inter=128):
    return SimpleNamespace(hidden_size=hidden, num_attention_heads=heads, intermediate_size=inter,
                           attention_probs_dropout_prob=0.0, hidden_dropout_prob=0.0, layer_norm_eps=1e-5,
                           is_decoder=False, position_embedding_type="absolute", max_position_embeddings=64)


def bert_config_hf(hidden=64, heads=4, inter=128):
    from transformers import BertConfig
    return BertConfig(hidden_size=hidden, num_attention_heads=heads, intermediate_size=inter, hidden_dropout_prob=0.0,
                      attention_probs_dropout_prob=0.0, layer_norm_eps=1e-5, max_position_embeddings=64)


class BertHostLayer(nn.Module):
    def __init__(self, blocks, cfg):
        super().__init__()
        self.attention = blocks.BertSelfAttention(cfg)
        self.attn_out = blocks.BertSelfOutput(cfg)
        self.inter = nn.Linear(cfg.hidden_size, cfg.intermediate_size)
        self.act = nn.GELU()
        self.out = blocks.BertOutput(cfg)

    def forward(self, x, mask):
        a = self.attention(x, mask)[0]
        a = self.attn_out(a, x)
        return self.out(self.act(self.inter(a)), a)


class BertHost(nn.Module):
    """RoBERTa-style classifier: encoder layers, then the first token through dense -> tanh -> out_proj (the head
    the GLUE recipe fine-tunes; BASELINE configs[3] leaves it quantized)."""

    def __init__(self, blocks, cfg, layers=2, labels=3):
        super().__init__()
        self.layers = nn.ModuleList([BertHostLayer(blocks, cfg) for _ in range(layers)])
        self.dense = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.out_proj = nn.Linear(cfg.hidden_size, labels)

    def forward(self, x, mask):
        for layer in self.layers:
            x = layer(x, mask)
        return self.out_proj(torch.tanh(self.dense(x[:, 0])))


# ----------------------------------------------------------------------------------------------- Llama
def llama_config(hidden=128, heads=4, inter=256, vocab=128, layers=2, max_pos=128):
    from transformers import LlamaConfig
    cfg = LlamaConfig(hidden_size=hidden, num_attention_heads=heads, num_key_value_heads=heads,
                      intermediate_size=inter, vocab_size=vocab, num_hidden_layers=layers,
                      max_position_embeddings=max_pos, rms_norm_eps=1e-5, attention_dropout=0.0,
                      attention_bias=False, mlp_bias=False, hidden_act="silu")
    # attributes the reference's transformers-4.36-era block reads (modeling_llama.py:110-160)
    for name, value in (("pretraining_tp", 1), ("rope_theta", 10000.0), ("rope_scaling", None)):
        try:
            if getattr(cfg, name, None) != value:
                setattr(cfg, name, value)
        except Exception:
            object.__setattr__(cfg, name, value)
    return cfg


def rope_tables(head_dim, seq_len, base=10000.0, dtype=torch.bfloat16, device=None):
    """cos / sin [seq_len, head_dim] exactly as transformers 4.36's LlamaRotaryEmbedding caches them (fp32 outer
    product, cat, cos / sin, then cast) -- and as transformers 5.x computes them per call."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    t = torch.arange(seq_len, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype).to(device), emb.sin().to(dtype).to(device)


def causal_mask(batch, seq, dtype=torch.bfloat16, device=None):
    m = torch.full((seq, seq), torch.finfo(dtype).min, dtype=dtype)
    m = torch.triu(m, diagonal=1)
    return m[None, None].expand(batch, 1, seq, seq).contiguous().to(device)


class LlamaHost(nn.Module):
    """embed -> decoder layers -> final RMSNorm -> lm_head.  `call_layer(layer, x, mask, position_ids)` adapts the
    call signature (reference: transformers 4.36 keywords, returns a tuple; this repo: transformers 5.x)."""

    def __init__(self, blocks, cfg, call_layer):
        super().__init__()
        from transformers.models.llama.modeling_llama import LlamaRMSNorm
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList([blocks.LlamaDecoderLayer(cfg, i) for i in range(cfg.num_hidden_layers)])
        self.norm = LlamaRMSNorm(cfg.hidden_size, eps=cfg.rms_norm_eps)
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        self._call_layer = call_layer

    def forward(self, input_ids, mask):
        x = self.embed_tokens(input_ids)
        position_ids = torch.arange(input_ids.shape[1], device=input_ids.device)[None].expand(input_ids.shape[0], -1)
        for layer in self.layers:
            x = self._call_layer(layer, x, mask, position_ids)
        return self.lm_head(self.norm(x))


def nll(logits, input_ids):
    """HF causal-LM loss (labels = input_ids, shifted), fp32: the quantity behind the reference's perplexity
    (examples/language_modeling/wikitext.py:146-167)."""
    lg = logits[:, :-1].float().reshape(-1, logits.shape[-1])
    return torch.nn.functional.cross_entropy(lg, input_ids[:, 1:].reshape(-1))


def state_to_numpy(model, prefix="w/"):
    return {prefix + k: v.detach().float().numpy().copy() for k, v in model.state_dict().items()
            if v.dtype.is_floating_point}
