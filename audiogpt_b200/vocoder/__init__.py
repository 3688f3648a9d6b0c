"""Mirror of text_to_audio/Make_An_Audio/vocoder/ (a namespace directory in the reference)."""
