// media_b200/csrc/rate_control.h -- frame-level rate control of the B200 H.264 encoder (host logic, plain C++).
//
// What it stands in for: the rate control openh264 runs for the wrapper's configuration -- RC_BITRATE_MODE with
// iMaxBitrate = iTargetBitrate (reference video_codec/VideoEncoderOpenH264.cpp:239-240,274) and the iMinQp / iMaxQp the wrapper
// takes from GetDefaultParams (:230; vendor/openh264/codec_app_def.h SEncParamExt). openh264's own model is not in the tree, so
// this is OUR specification: one QP per picture, chosen on the host before the picture is launched.
//
//  * target: T = bitrate / fps bits per picture; an IDR may take IDR_WEIGHT pictures' worth.
//  * feedback: a virtual buffer collects (bits - T) per picture; its level is paid back over half a second.
//  * model: per picture type log2(bits) = a - alpha * qp / 6. Both a and alpha are learnt online (alpha from consecutive
//    pictures coded with different QPs; this encoder's alpha is 0.3 .. 1.0 depending on content and QP range, so the
//    textbook alpha = 1 overshoots every correction).
//  * bounds: QP always inside [min_qp, max_qp] of the configuration (iMinQp / iMaxQp), moves at most +-3 between P pictures.
//  * max bitrate: a leaky bucket of one second drained at max_bitrate (VBV); a picture's budget never exceeds the room left,
//    and a picture that comes out larger than its hard cap (IDR: IDR_CAP pictures' worth, P: P_CAP; never more than the room
//    in the bucket) is coded again once with a coarser QP before it is delivered (retry_qp).
// Compiled into libb200enc.so (engine.cu) and, for the CPU tests and tools/rc_sim.py, on its own (rate_control_capi.cpp).
#pragma once
#include <algorithm>
#include <cmath>

namespace b200rc {

struct Config {
    double bitrate = 0, max_bitrate = 0;   // bits per second; max_bitrate <= 0: same as bitrate (the wrapper's setting)
    int fps = 30, min_qp = 0, max_qp = 51;
    int width = 0, height = 0;
};
struct Decision {
    int qp = 30;
    double budget = 0;      // bits the model aimed at
    double hard_cap = 0;    // bits above which the picture is coded again (0: never)
};

class RateCtl {
public:
    static constexpr double IDR_WEIGHT = 5.0, IDR_CAP = 8.0, P_CAP = 6.0;
    static constexpr int QP_FLOOR = 10;     // the model never goes finer than this on its own (openh264 keeps camera content above 12)

    void init(const Config &c)
    {
        cfg = c;
        if (cfg.fps <= 0) cfg.fps = 30;
        if (cfg.max_bitrate <= 0 || cfg.max_bitrate < cfg.bitrate) cfg.max_bitrate = cfg.bitrate;
        cfg.min_qp = std::min(std::max(cfg.min_qp, 0), 51); cfg.max_qp = std::min(std::max(cfg.max_qp, cfg.min_qp), 51);
        T = cfg.bitrate / cfg.fps; drain = cfg.max_bitrate / cfg.fps; bucket = cfg.max_bitrate;
        level = 0; vbv = 0;
        m[0] = Model(); m[1] = Model();
        last_qp[0] = last_qp[1] = -1; prev_type = -1;
    }
    int lo() const { return std::max(cfg.min_qp, std::min(QP_FLOOR, cfg.max_qp)); }
    int hi() const { return cfg.max_qp; }

    Decision pick(int type) const
    {
        Decision d;
        const double w = type == 1 ? IDR_WEIGHT : 1.0, react = std::max(2.0, cfg.fps * 0.5);
        double want = T * w - level / react;
        want = std::min(std::max(want, (type == 1 ? 2.0 : 0.3) * T), (type == 1 ? IDR_WEIGHT : 2.5) * T);
        const double room = std::max(bucket - vbv, 0.25 * T);            // what the one-second bucket still takes
        want = std::min(want, std::max(room * 0.8, 0.25 * T));
        d.budget = want;
        d.hard_cap = std::max(std::min((type == 1 ? IDR_CAP : P_CAP) * T, room), 1.5 * want);
        int qp;
        if (m[type].have && type == 0) {
            // P pictures move in damped steps from the last P QP: a picture coded finer than its reference pays for the reference's
            // error as well (and the other way round), so the apparent exponent between alternating QPs is ~2 and full steps oscillate
            const double dq = m[0].dqp_for(want);
            const int step = std::fabs(dq) < 0.8 ? 0 : (int)std::lround(0.75 * dq);
            qp = last_qp[0] + std::min(std::max(step, -2), 3);
            // the first P picture behind a key picture predicts from it: coded much finer than the key picture it would pay for the whole
            // picture's quantisation error at once (measured: 12 QP steps finer = 10-40 picture budgets); it starts 3 steps below the key
            // picture's QP and the usual -2 per picture bring the sequence back
            if (prev_type == 1 && last_qp[1] >= 0) qp = std::max(qp, last_qp[1] - 3);
        } else if (m[type].have) {
            qp = m[type].qp_for(want);
            if (last_qp[0] >= 0) qp = std::min(std::max(qp, last_qp[0] - 3), last_qp[0] + 8);
            // ... but never a QP at which the key-frame model itself predicts a picture above the hard cap: that attempt would be thrown away and
            // coded again (content whose P pictures are cheap and whose key pictures are not -- a translating texture -- sits 15+ steps apart)
            while (qp < hi() && m[type].predict_l2(qp) > std::log2(0.9 * d.hard_cap)) qp++;
        } else if (type == 0 && last_qp[1] >= 0) {
            qp = last_qp[1] - 3;                                         // first P picture after the first IDR
        } else if (type == 1 && m[0].have) {
            qp = last_qp[0] + 2;                                         // scene-change IDR before any IDR statistics: start from the P QP
        } else {
            const double bpp = cfg.width > 0 ? cfg.bitrate / cfg.fps / ((double)cfg.width * cfg.height) : 0.05;
            // bits per pixel per picture -> a starting QP (an intra picture costs ~10x a P picture at equal QP: start it coarser)
            qp = bpp > 0.2 ? 24 : bpp > 0.1 ? 28 : bpp > 0.05 ? 32 : bpp > 0.02 ? 36 : 40;
            if (type == 1) qp += 5;
        }
        d.qp = std::min(std::max(qp, lo()), hi());
        return d;
    }
    // QP for a second attempt when `bits` came out above the hard cap of decision d, or -1 to accept the picture
    int retry_qp(int type, const Decision &d, double bits) const
    {
        if (d.hard_cap <= 0 || bits <= d.hard_cap || d.qp >= hi()) return -1;
        const double alpha = m[type].have ? m[type].alpha : 0.6;
        const double aim = std::max(d.budget, 0.6 * d.hard_cap);
        const int up = (int)std::ceil(6.0 * std::log2(bits / aim) / alpha);
        return std::min(d.qp + std::min(std::max(up, 2), 12), hi());
    }
    // The whole second-attempt rule, shared by the engine and the simulation: `planned` is the picture type the QP was picked for (decision
    // d), `coded` what it came out as (a P picture the device promoted to a scene-change IDR is judged against an IDR's cap).
    int second_attempt_qp(int planned, int coded, const Decision &d, double bits)
    {
        Decision e = d;
        if (coded && !planned) { e = pick(1); e.qp = d.qp; }
        const int q2 = retry_qp(coded, e, bits);
        if (q2 >= 0) note_discarded(coded, d.qp, bits);      // the attempt is thrown away, what it measured is not
        return q2;
    }
    // A first attempt that was thrown away (coded again at another QP) still measured the picture: the model learns from it -- two points
    // of the same picture give the exponent alpha at once -- but the buffers only ever see the delivered picture.
    void note_discarded(int type, int qp, double bits) { m[type].observe(qp, std::max(bits, 64.0)); }
    void update(int type, int qp, double bits)
    {
        bits = std::max(bits, 64.0);
        m[type].observe(qp, bits);
        last_qp[type] = qp; prev_type = type;
        level += bits - T;
        level = std::min(std::max(level, -1.0 * cfg.fps * T), 4.0 * cfg.fps * T);
        vbv = std::max(0.0, vbv + bits - drain);
    }
    double vbv_level() const { return vbv; }
    double vbv_size() const { return bucket; }
    double buffer_level() const { return level; }
    double alpha(int type) const { return m[type].alpha; }

private:
    struct Model {
        bool have = false; double lbits = 0; int qp = 30; double alpha = 0.6;
        double predict_l2(int q) const { return lbits - alpha * (q - qp) / 6.0; }
        double dqp_for(double want) const { return 6.0 * (lbits - std::log2(want)) / alpha; }
        int qp_for(double want) const { return qp + (int)std::lround(dqp_for(want)); }
        void observe(int q, double bits)
        {
            const double lb = std::log2(bits);
            if (have && q != qp) {
                const double a = (lbits - lb) / ((q - qp) / 6.0);
                if (a > 0.05 && a < 4.0) alpha = 0.6 * alpha + 0.4 * std::min(std::max(a, 0.25), 2.0);
                lbits = 0.3 * predict_l2(q) + 0.7 * lb;                // the old estimate carried to the new QP, blended with the observation
            } else lbits = have ? 0.5 * lbits + 0.5 * lb : lb;
            qp = q; have = true;
        }
    };
    Config cfg; double T = 0, drain = 0, bucket = 0, level = 0, vbv = 0;
    Model m[2]; int last_qp[2] = { -1, -1 }, prev_type = -1;
};

} // namespace b200rc
