// nl_sample.cuh — device-side sampling step (nl_sample.cu): repetition penalty + sampleTopK / sampleTopP / argmax of go/main.go.
#pragma once
#include "nl_common.cuh"

namespace nl {

struct SampleArgs {
    float *logits;            // [vocab] logits of the last forward (the repetition penalty is applied in place, like the reference)
    int vocab;
    const int32_t *recent;    // device copy of the repetition window
    int n_recent;
    float rep_penalty;
    float temp;
    int top_k;
    float top_p;
    float u;                  // rng.Float32() drawn by the host
    uint32_t *keys0, *keys1;  // sort scratch, [vocab] each
    int32_t *idx0, *idx1;
    int32_t *token_out;       // device scalar
};

int launch_sample(const SampleArgs &a, cudaStream_t st);

}  // namespace nl
