/* Shim: utils encode.h is included by lib/aln/bntseq.c but its only use there is commented
 * out (bntseq.c:581). Intentionally empty. TEST INFRASTRUCTURE ONLY. */
