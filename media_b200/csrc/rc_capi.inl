// media_b200/csrc/rc_capi.inl -- C entry points of the rate control (rate_control.h) for the CPU tests and tools/rc_sim.py: the
// same object the sessions use, driven from outside with picture sizes (no GPU involved). Included by engine.cu.
struct B200RcBox { b200rc::RateCtl rc; b200rc::Decision last; };
extern "C" {
void *b200k_rc_create(double bitrate, double max_bitrate, int fps, int min_qp, int max_qp, int width, int height)
{
    B200RcBox *b = new (std::nothrow) B200RcBox();
    if (!b) return nullptr;
    b200rc::Config c; c.bitrate = bitrate; c.max_bitrate = max_bitrate; c.fps = fps; c.min_qp = min_qp; c.max_qp = max_qp; c.width = width; c.height = height;
    b->rc.init(c);
    return b;
}
int b200k_rc_pick(void *h, int type, double *budget, double *hard_cap)
{
    B200RcBox *b = static_cast<B200RcBox *>(h);
    b->last = b->rc.pick(type);
    if (budget) *budget = b->last.budget;
    if (hard_cap) *hard_cap = b->last.hard_cap;
    return b->last.qp;
}
int b200k_rc_retry_qp(void *h, int planned_type, int coded_type, double bits)
{
    B200RcBox *b = static_cast<B200RcBox *>(h);
    return b->rc.second_attempt_qp(planned_type, coded_type, b->last, bits);
}
void b200k_rc_update(void *h, int type, int qp, double bits) { static_cast<B200RcBox *>(h)->rc.update(type, qp, bits); }
double b200k_rc_vbv(void *h, double *size) { B200RcBox *b = static_cast<B200RcBox *>(h); if (size) *size = b->rc.vbv_size(); return b->rc.vbv_level(); }
void b200k_rc_destroy(void *h) { delete static_cast<B200RcBox *>(h); }
}
