"""Drop-in mirror of the part of the reference's `longva` package the streaming driver uses
(inference_streaming_longva_v2.py:11-17,697): builder.load_pretrained_model, the model object's
encode_images / generate_with_image_embedding / get_model().embed_tokens, the qwen conversation
templates, tokenizer_image_token and the constants.  `dropin/longva` re-exports this package under
the reference's import paths."""
