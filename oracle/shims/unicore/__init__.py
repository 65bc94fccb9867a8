"""CPU restatement of the slice of Uni-Core (github.com/dptech-corp/Uni-Core, unpinned,
not vendored by MM-DTI) that the reference imports:
  models/transformers.py:11   unicore.modules.{TransformerEncoderLayer, LayerNorm}
  models/mm_model.py:13-16    unicore.utils.get_activation_fn, unicore.data.Dictionary,
                              unicore.models.BaseUnicoreModel, unicore.modules.init_bert_params
TEST INFRASTRUCTURE: parity for this part is UNPINNED (see oracle/__init__.py)."""
