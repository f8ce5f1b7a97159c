/* The reference tree omits libsamplerate's large high-quality coefficient blob.
 * The mid-quality table is included first by src_sinc.c, so alias to it.  Resampler
 * quality is irrelevant to decode parity (parity is defined on stream bytes -> PCM). */
#define slow_high_qual_coeffs slow_mid_qual_coeffs
