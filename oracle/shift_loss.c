/* Oracle (plain C, double precision): shift-compensated scores.  TEST INFRASTRUCTURE ONLY.
 *
 * Second, independent restatement of reference models/loss.py used to cross-check
 * oracle/losses.py and as a fast CPU baseline at large batch:
 *   stackL1Loss :140-152, stackL2Loss :154-166, stackcPSNR :168-180, stackL1EdgeLoss :126-138,
 *   computeBiasBrightness :182-187, computeL1EdgeLoss :219-224, computeL1Loss :226-228,
 *   computeL2Loss :230-232, computecPSNR :234-238, cropImage utils/utils.py:42-44.
 * Layouts: hr, sr float32 [B,H,W]; mask uint8 [B,H,W] (1 = clear).
 * scores[b*S*S + i*S + j] with S = 2*border+1 (reference stack order, loss.py:48-50).
 * PARITY UNPINNED (no TensorFlow in the image).
 *
 * kind: 0 = L1, 1 = L2, 2 = cPSNR, 3 = L1Edge (0.7*L1 + 0.3*sobel)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline int reflect(int i, int n) { /* tf.pad REFLECT by 1 */
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

void oracle_shift_scores(int kind, const float* hr, const uint8_t* mask, const float* sr,
                         int B, int H, int W, int border, double* scores, double* counts, double* biases) {
    const int S = 2 * border + 1;
    const int ch = H - 2 * border, cw = W - 2 * border;
    double* r = (double*)malloc(sizeof(double) * (size_t)ch * cw);
    for (int b = 0; b < B; ++b) {
        const float* h_ = hr + (size_t)b * H * W;
        const float* p_ = sr + (size_t)b * H * W;
        const uint8_t* m_ = mask + (size_t)b * H * W;
        for (int i = 0; i < S; ++i)
            for (int j = 0; j < S; ++j) {
                double N = 0, sh = 0, spm = 0;
                for (int y = 0; y < ch; ++y)
                    for (int x = 0; x < cw; ++x) {
                        double h = h_[(y + i) * W + (x + j)];
                        double m = m_[(y + i) * W + (x + j)] ? 1.0 : 0.0;
                        double p = p_[(y + border) * W + (x + border)];
                        N += m; sh += h; spm += p * m;
                    }
                double bias = (sh - spm) / N;
                double a1 = 0, a2 = 0;
                for (int y = 0; y < ch; ++y)
                    for (int x = 0; x < cw; ++x) {
                        double h = h_[(y + i) * W + (x + j)];
                        double m = m_[(y + i) * W + (x + j)] ? 1.0 : 0.0;
                        double p = p_[(y + border) * W + (x + border)];
                        double d = h - (p + bias) * m;
                        r[y * cw + x] = d;
                        a1 += fabs(d); a2 += d * d;
                    }
                double sc;
                if (kind == 0) sc = a1 / N;
                else if (kind == 1) sc = a2 / N;
                else if (kind == 2) sc = 10.0 * (log(65535.0 * 65535.0 / (a2 / N)) / log(10.0));
                else {
                    /* sobel(h) - sobel(corr) = sobel(h - corr) (linear; both reflect-padded on the crop) */
                    double e = 0;
                    for (int y = 0; y < ch; ++y)
                        for (int x = 0; x < cw; ++x) {
                            int ym = reflect(y - 1, ch), yp = reflect(y + 1, ch);
                            int xm = reflect(x - 1, cw), xp = reflect(x + 1, cw);
                            double gy = -r[ym * cw + xm] - 2 * r[ym * cw + x] - r[ym * cw + xp]
                                        + r[yp * cw + xm] + 2 * r[yp * cw + x] + r[yp * cw + xp];
                            double gx = -r[ym * cw + xm] + r[ym * cw + xp] - 2 * r[y * cw + xm]
                                        + 2 * r[y * cw + xp] - r[yp * cw + xm] + r[yp * cw + xp];
                            e += fabs(gy) + fabs(gx);
                        }
                    sc = 0.7 * (a1 / N) + (1.0 - 0.7) * (e / N);
                }
                size_t o = (size_t)b * S * S + (size_t)i * S + j;
                scores[o] = sc;
                if (counts) counts[o] = N;
                if (biases) biases[o] = bias;
            }
    }
    free(r);
}
