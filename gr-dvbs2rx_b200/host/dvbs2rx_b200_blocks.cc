// dvbs2rx_b200_blocks.cc -- see dvbs2rx_b200_blocks.h.  Pure host C++ above include/dvbs2_b200.h.
#include "dvbs2rx_b200_blocks.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>

#include "../../include/dvbs2_b200.h"

namespace gr {
namespace dvbs2rx {

namespace {
// Device the blocks of this process are created on: set_device(), else the DVBS2RX_B200_DEVICE environment variable,
// else 0.  Every block owns its handle (tables, streams, scratch): GNU Radio runs the blocks of a flowgraph in
// different threads and a handle serves one thread at a time; sharing one between the demapper, LDPC and BCH blocks
// would serialise them behind a lock for the sake of two table uploads of ~270 KB.
int g_device = -1;
int block_device()
{
    if (g_device >= 0)
        return g_device;
    const char* env = getenv("DVBS2RX_B200_DEVICE");
    return env ? atoi(env) : 0;
}
dvbs2b200_code* create_code(int standard, int framesize, int rate)
{
    dvbs2b200_code* h = nullptr;
    int rc = dvbs2b200_code_create(&h, block_device(), standard, framesize, rate);
    if (rc != DVBS2B200_OK)
        throw std::runtime_error(std::string("dvbs2b200: ") + dvbs2b200_last_error());
    return h;
}
const int DEFAULT_TRIALS = 25; // lib/ldpc_decoder_bb_impl.cc:391
} // namespace

void set_device(int device) { g_device = device; }

// ---- ldpc_decoder_bb ------------------------------------------------------------------------------
ldpc_decoder_bb::sptr ldpc_decoder_bb::make(dvb_standard_t standard, dvb_framesize_t framesize, dvb_code_rate_t rate,
                                            dvb_constellation_t /*constellation*/, dvb_outputmode_t outputmode,
                                            dvb_infomode_t /*infomode*/, int max_trials, int /*debug_level*/)
{
    sptr b(new ldpc_decoder_bb());
    b->d_code = create_code(standard, framesize, rate);
    dvbs2b200_code_info info;
    dvbs2b200_code_info_get(b->d_code, &info);
    b->d_kldpc = info.nbch; // lib/ldpc_decoder_bb_impl.cc:97-102: kldpc = bch.n
    b->d_nldpc = info.n_ldpc;
    b->d_kldpc_bytes = b->d_kldpc / 8;
    b->d_nldpc_bytes = b->d_nldpc / 8;
    b->d_output_mode = outputmode;
    b->d_max_trials = max_trials;
    b->set_batches_per_call(1);
    return b;
}

ldpc_decoder_bb::~ldpc_decoder_bb() { dvbs2b200_code_destroy(d_code); }

void ldpc_decoder_bb::set_batches_per_call(int n)
{
    // lib/ldpc_decoder_bb_impl.cc:354-360 with simd -> n * 32 frames per scheduler call
    const int frames = std::max(1, n) * d_simd_size;
    if (d_output_mode == OM_MESSAGE) {
        d_output_multiple = d_kldpc_bytes * frames;
        d_relative_rate = (double)d_kldpc_bytes / d_nldpc;
    } else {
        d_output_multiple = d_nldpc_bytes * frames;
        d_relative_rate = (double)d_nldpc_bytes / d_nldpc;
    }
}

void ldpc_decoder_bb::forecast(int noutput_items, gr_vector_int& ninput_items_required)
{
    if (d_output_mode == OM_MESSAGE) { // lib/ldpc_decoder_bb_impl.cc:380-389
        unsigned int n_frames = noutput_items / d_kldpc_bytes;
        ninput_items_required[0] = n_frames * d_nldpc;
    } else {
        ninput_items_required[0] = 8 * noutput_items;
    }
}

int ldpc_decoder_bb::general_work(int noutput_items, gr_vector_int& /*ninput_items*/,
                                  gr_vector_const_void_star& input_items, gr_vector_void_star& output_items)
{
    const int8_t* in = (const int8_t*)input_items[0];
    unsigned char* out = (unsigned char*)output_items[0];
    const int trials = (d_max_trials == 0) ? DEFAULT_TRIALS : d_max_trials;
    const int output_size = d_output_mode ? d_kldpc_bytes : d_nldpc_bytes;
    const int frames = noutput_items / output_size; // a multiple of d_simd_size by the output multiple
    if (frames <= 0 || frames % d_simd_size) {
        d_consumed = 0;
        return 0;
    }
    const bool want_pdu = (bool)d_pdu_handler;
    if (want_pdu)
        d_post.resize((size_t)frames * d_nldpc);
    d_ret.resize(frames);
    // one launch for the whole call; groups of 32 frames keep the reference's coupled iteration loop
    int rc = dvbs2b200_ldpc_decode(d_code, in, frames, trials, d_simd_size, d_output_mode, out,
                                   want_pdu ? d_post.data() : nullptr, d_ret.data());
    if (rc != DVBS2B200_OK)
        throw std::runtime_error(std::string("dvbs2b200_ldpc_decode: ") + dvbs2b200_last_error());
    for (int b = 0; b < frames / d_simd_size; ++b) {
        const int count = d_ret[(size_t)b * d_simd_size];
        d_total_trials += (count < 0) ? trials : (trials - count); // lib/ldpc_decoder_bb_impl.cc:411-419
        if (want_pdu) {
            llr_pdu pdu;
            pdu.simd_size = d_simd_size;
            pdu.frame_cnt = d_frame_cnt;
            pdu.llr = d_post.data() + (size_t)b * d_simd_size * d_nldpc;
            pdu.n_llr = (size_t)d_simd_size * d_nldpc;
            d_pdu_handler(pdu);
        }
        d_frame_cnt += d_simd_size;
        d_batch_cnt++;
    }
    d_consumed = frames * d_nldpc; // consume_each
    return frames * output_size;
}

// ---- bch_decoder_bb -------------------------------------------------------------------------------
bch_decoder_bb::sptr bch_decoder_bb::make(dvb_standard_t standard, dvb_framesize_t framesize, dvb_code_rate_t rate,
                                          dvb_outputmode_t /*outputmode*/, int /*debug_level*/)
{
    sptr b(new bch_decoder_bb());
    b->d_code = create_code(standard, framesize, rate);
    dvbs2b200_code_info info;
    dvbs2b200_code_info_get(b->d_code, &info);
    if (info.kbch % 8 || info.nbch % 8) // lib/bch.cc:19-24 assert_byte_aligned_n_k
        throw std::runtime_error("u8 array messages are only supported for n and k multiple of 8.");
    b->d_k_bytes = info.kbch / 8;
    b->d_n_bytes = info.nbch / 8;
    return b;
}

bch_decoder_bb::~bch_decoder_bb() { dvbs2b200_code_destroy(d_code); }

void bch_decoder_bb::forecast(int noutput_items, gr_vector_int& ninput_items_required)
{
    ninput_items_required[0] = (noutput_items / d_k_bytes) * d_n_bytes; // lib/bch_decoder_bb_impl.cc:78-82
}

int bch_decoder_bb::general_work(int noutput_items, gr_vector_int& /*ninput_items*/,
                                 gr_vector_const_void_star& input_items, gr_vector_void_star& output_items)
{
    const unsigned char* in = (const unsigned char*)input_items[0];
    unsigned char* out = (unsigned char*)output_items[0];
    const int n_codewords = noutput_items / d_k_bytes;
    d_corr.resize(std::max(n_codewords, 1));
    int rc = dvbs2b200_bch_decode(d_code, in, n_codewords, out, d_corr.data());
    if (rc != DVBS2B200_OK)
        throw std::runtime_error(std::string("dvbs2b200_bch_decode: ") + dvbs2b200_last_error());
    for (int i = 0; i < n_codewords; i++) { // lib/bch_decoder_bb_impl.cc:94-111
        if (d_corr[i] == -1)
            d_frame_error_cnt++;
        d_frame_cnt++;
    }
    d_consumed = n_codewords * d_n_bytes;
    return noutput_items;
}

// ---- bbdescrambler_bb -----------------------------------------------------------------------------
bbdescrambler_bb::sptr bbdescrambler_bb::make(dvb_standard_t standard, dvb_framesize_t framesize, dvb_code_rate_t rate)
{
    sptr b(new bbdescrambler_bb());
    b->d_code = create_code(standard, framesize, rate);
    dvbs2b200_code_info info;
    dvbs2b200_code_info_get(b->d_code, &info);
    b->kbch_bytes = info.kbch / 8; // lib/bbdescrambler_bb_impl.cc:41-44
    return b;
}

bbdescrambler_bb::~bbdescrambler_bb() { dvbs2b200_code_destroy(d_code); }

int bbdescrambler_bb::work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items)
{
    // lib/bbdescrambler_bb_impl.cc:67-82; noutput_items is a multiple of kbch_bytes (set_output_multiple)
    int rc = dvbs2b200_bb_descramble(d_code, (const unsigned char*)input_items[0], noutput_items / (int)kbch_bytes,
                                     (unsigned char*)output_items[0]);
    if (rc != DVBS2B200_OK)
        throw std::runtime_error(std::string("dvbs2b200_bb_descramble: ") + dvbs2b200_last_error());
    return noutput_items;
}

// ---- bbdeheader_bb --------------------------------------------------------------------------------
bbdeheader_bb::sptr bbdeheader_bb::make(dvb_standard_t standard, dvb_framesize_t framesize, dvb_code_rate_t rate,
                                        int /*debug_level*/)
{
    sptr b(new bbdeheader_bb());
    b->d_code = create_code(standard, framesize, rate);
    dvbs2b200_code_info info;
    dvbs2b200_code_info_get(b->d_code, &info);
    b->d_kbch_bytes = info.kbch / 8;   // lib/bbdeheader_bb_impl.cc:57-61
    b->d_max_dfl = info.kbch - 80;
    return b;
}

bbdeheader_bb::~bbdeheader_bb() { dvbs2b200_code_destroy(d_code); }

void bbdeheader_bb::forecast(int noutput_items, gr_vector_int& ninput_items_required)
{
    // lib/bbdeheader_bb_impl.cc:69-74
    unsigned int n_bbframes = (unsigned int)std::ceil(static_cast<double>(noutput_items * 8) / d_max_dfl);
    ninput_items_required[0] = n_bbframes * d_kbch_bytes;
}

int bbdeheader_bb::general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                                gr_vector_void_star& output_items)
{
    // lib/bbdeheader_bb_impl.cc:154-160: as many whole BBFRAMEs as are available and fit
    const unsigned int in_bbframes = ninput_items[0] / d_kbch_bytes;
    const unsigned int out_bbframes = (unsigned int)std::ceil(static_cast<double>(noutput_items * 8) / d_max_dfl);
    const unsigned int n_bbframes = std::min(in_bbframes, out_bbframes);
    size_t produced = 0;
    // like the reference, up to one packet more than noutput_items can come out (a carried partial packet)
    int rc = dvbs2b200_bb_deheader(d_code, (const unsigned char*)input_items[0], (int)n_bbframes, d_scrambled ? 1 : 0,
                                   (unsigned char*)output_items[0], dvbs2b200_bb_ts_capacity(d_code, (int)n_bbframes), &produced);
    if (rc != DVBS2B200_OK)
        throw std::runtime_error(std::string("dvbs2b200_bb_deheader: ") + dvbs2b200_last_error());
    d_consumed = (int)(n_bbframes * d_kbch_bytes);
    return (int)produced;
}

uint64_t bbdeheader_bb::counters(int which)
{
    dvbs2b200_bb_counters c;
    if (dvbs2b200_bb_counters_get(d_code, &c) != DVBS2B200_OK)
        throw std::runtime_error(std::string("dvbs2b200_bb_counters_get: ") + dvbs2b200_last_error());
    const uint64_t v[5] = { c.packets, c.errors, c.bbframes, c.dropped, c.gaps };
    return v[which];
}

// ---- xfecframe_demapper_cb ------------------------------------------------------------------------
xfecframe_demapper_cb::sptr xfecframe_demapper_cb::make(dvb_framesize_t framesize, dvb_code_rate_t rate,
                                                        dvb_constellation_t constellation, bool apsk_opt_in)
{
    sptr b(new xfecframe_demapper_cb());
    b->d_constellation = constellation;
    b->d_rate = rate;
    if (framesize == FECFRAME_NORMAL)
        b->d_fecframe_len = 64800;
    else if (framesize == FECFRAME_MEDIUM)
        b->d_fecframe_len = 32400;
    else
        b->d_fecframe_len = 16200;
    if (constellation == MOD_QPSK) {
        b->d_bits = 2;
    } else if (constellation == MOD_8PSK) {
        b->d_bits = 3;
        unsigned int rows = b->d_fecframe_len / 3; // lib/xfecframe_demapper_cb_impl.cc:48-69
        if (rate == C3_5) {
            b->d_rowaddr0 = rows * 2, b->d_rowaddr1 = rows, b->d_rowaddr2 = 0;
        } else if (rate == C25_36 || rate == C13_18 || rate == C7_15 || rate == C8_15 || rate == C26_45) {
            b->d_rowaddr0 = rows, b->d_rowaddr1 = 0, b->d_rowaddr2 = rows * 2;
        } else {
            b->d_rowaddr0 = 0, b->d_rowaddr1 = rows, b->d_rowaddr2 = rows * 2;
        }
    } else if (apsk_opt_in && (constellation == MOD_16APSK || constellation == MOD_32APSK) &&
               apsk_points(constellation, rate, b->d_points)) {
        // not in the reference: bit k of symbol j sits in interleaver column k (EN 302 307-1 5.3.3), rows = N / bits
        b->d_bits = constellation == MOD_16APSK ? 4 : 5;
        b->d_table = true;
        b->d_waiting_first_llr = false; // no estimator for APSK: N0 comes from set_es_n0_db()
        for (unsigned int k = 0; k < b->d_bits; ++k)
            b->d_row_offsets.push_back((int)(k * (b->d_fecframe_len / b->d_bits)));
    } else {
        throw std::runtime_error("Unsupported constellation");
    }
    b->d_xfecframe_len = b->d_fecframe_len / b->d_bits;
    for (size_t i = 0; i < kPool; i++) {
        b->d_saved[i] = std::numeric_limits<uint64_t>::max();
        b->d_pool[i].resize(b->d_xfecframe_len);
    }
    b->d_code = create_code(STANDARD_DVBS2, framesize, rate);
    return b;
}

xfecframe_demapper_cb::~xfecframe_demapper_cb() { dvbs2b200_code_destroy(d_code); }

void xfecframe_demapper_cb::set_es_n0_db(float es_n0_db)
{
    std::lock_guard<std::mutex> l(d_mutex);
    d_snr = es_n0_db;
    d_N0 = std::pow(10.0f, -es_n0_db / 10.0f);
    d_precision = 4.0f / d_N0;
}

void xfecframe_demapper_cb::forecast(int noutput_items, gr_vector_int& ninput_items_required)
{
    ninput_items_required[0] = noutput_items / d_bits;
}


// The SNR estimates (lib/qpsk.h:240-281, lib/xfecframe_demapper_cb_impl.cc:128-142,267-302) run on the GPU:
// dvbs2b200_estimate_snr, one launch per general_work call / per PDU.
int xfecframe_demapper_cb::general_work(int noutput_items, gr_vector_int& /*ninput_items*/,
                                        gr_vector_const_void_star& input_items, gr_vector_void_star& output_items)
{
    std::lock_guard<std::mutex> l(d_mutex);
    const gr_complex* in = static_cast<const gr_complex*>(input_items[0]);
    int8_t* out = static_cast<int8_t*>(output_items[0]);
    const int n_frames = noutput_items / d_fecframe_len;
    d_n0_per_frame.resize(std::max(n_frames, 1));
    const gr_complex* p = in;
    // initial SNR estimates of the whole call in one launch (lib/xfecframe_demapper_cb_impl.cc:123-146)
    if (d_waiting_first_llr && n_frames > 0) {
        d_snr_per_frame.resize(n_frames);
        int rc = dvbs2b200_estimate_snr(d_code, d_constellation, reinterpret_cast<const float*>(in), nullptr, n_frames,
                                        d_snr_per_frame.data());
        if (rc != DVBS2B200_OK)
            throw std::runtime_error(std::string("dvbs2b200_estimate_snr: ") + dvbs2b200_last_error());
    }
    for (int i = 0; i < n_frames; i++) { // lib/xfecframe_demapper_cb_impl.cc:115-149
        d_saved[d_idx] = d_frame_cnt;
        memcpy(d_pool[d_idx].data(), p, d_xfecframe_len * sizeof(gr_complex));
        d_idx = (d_idx + 1) % kPool;
        if (d_waiting_first_llr) {
            float snr_lin = d_snr_per_frame[i];
            d_snr = 10 * std::log10(snr_lin);
            d_N0 = 1.0f / snr_lin;
            d_precision = 4.0 / d_N0;
        }
        d_n0_per_frame[i] = d_N0;
        p += d_xfecframe_len;
        d_frame_cnt++;
    }
    // soft demap + deinterleave of the whole call in one launch, N0 explicit per frame
    int rc = d_table ? dvbs2b200_demap_table(d_code, (int)d_bits, d_points.data(), d_row_offsets.data(), reinterpret_cast<const float*>(in),
                                             n_frames, d_n0_per_frame.data(), out)
                     : dvbs2b200_demap(d_code, d_constellation, reinterpret_cast<const float*>(in), n_frames, d_n0_per_frame.data(), out);
    if (rc != DVBS2B200_OK)
        throw std::runtime_error(std::string("dvbs2b200_demap: ") + dvbs2b200_last_error());
    d_consumed = n_frames * d_xfecframe_len;
    return noutput_items;
}

void xfecframe_demapper_cb::handle_llr_pdu(const llr_pdu& pdu)
{
    std::lock_guard<std::mutex> l(d_mutex);
    if (d_table)
        return; // no SNR refinement for the opt-in APSK constellations
    if (!pdu.llr || pdu.n_llr == 0 || pdu.n_llr != (size_t)pdu.simd_size * d_fecframe_len)
        return; // the reference logs and drops malformed PDUs (:193-242)
    size_t n_frames = pdu.n_llr / d_fecframe_len, n_processed = 0;
    float accum = 0;
    // gather the saved XFECFRAMEs of the PDU's frames, then one launch for all of them (:252-310)
    d_gather_iq.clear();
    d_gather_llr.clear();
    for (size_t i = 0; i < n_frames; i++) {
        size_t idx = kPool;
        for (size_t k = 0; k < kPool; k++)
            if (d_saved[k] == pdu.frame_cnt + i) {
                idx = k;
                break;
            }
        if (idx == kPool)
            continue;
        d_gather_iq.insert(d_gather_iq.end(), d_pool[idx].begin(), d_pool[idx].end());
        d_gather_llr.insert(d_gather_llr.end(), pdu.llr + i * d_fecframe_len, pdu.llr + (i + 1) * d_fecframe_len);
        n_processed++;
    }
    if (n_processed > 0) {
        d_snr_per_frame.resize(n_processed);
        int rc = dvbs2b200_estimate_snr(d_code, d_constellation, reinterpret_cast<const float*>(d_gather_iq.data()),
                                        d_gather_llr.data(), (int)n_processed, d_snr_per_frame.data());
        if (rc != DVBS2B200_OK)
            throw std::runtime_error(std::string("dvbs2b200_estimate_snr: ") + dvbs2b200_last_error());
        for (size_t i = 0; i < n_processed; i++)
            accum += d_snr_per_frame[i];
    }
    float avg = accum / n_processed; // :312-317 (NaN when nothing was processed, as the reference)
    d_snr = 10 * std::log10(avg);
    d_N0 = 1.0f / avg;
    d_precision = 4.0 / d_N0;
    if (n_processed > 0)
        d_waiting_first_llr = false;
}

} // namespace dvbs2rx
} // namespace gr

// ---- ldpc_cuda: the decode seam --------------------------------------------------------------------
// ---- EN 302 307-1 16APSK (clause 5.4.3) / 32APSK (clause 5.4.4) constellations -------------------------------
// The C++ twin of dvbs2rx_b200/apsk.py (same tables, checked against each other and for the structural invariants
// in tests/test_apsk_tables.py).  Nothing in the reference can pin them: parity UNPINNED.
bool apsk_points(gr::dvbs2rx::dvb_constellation_t constellation, gr::dvbs2rx::dvb_code_rate_t rate, std::vector<float>& points)
{
    using namespace gr::dvbs2rx;
    const double pi = 3.14159265358979323846;
    auto put = [&](double r, double ph) {
        points.push_back((float)(r * std::cos(ph)));
        points.push_back((float)(r * std::sin(ph)));
    };
    points.clear();
    if (constellation == MOD_16APSK) {
        double g;
        switch (rate) {
        case C2_3: g = 3.15; break;
        case C3_4: g = 2.85; break;
        case C4_5: g = 2.75; break;
        case C5_6: g = 2.70; break;
        case C8_9: g = 2.60; break;
        case C9_10: g = 2.57; break;
        default: return false;
        }
        const double r1 = std::sqrt(4.0 / (1.0 + 3.0 * g * g)), r2 = g * r1;
        const double outer[12] = { pi / 4, -pi / 4, 3 * pi / 4, -3 * pi / 4, pi / 12, -pi / 12, 11 * pi / 12, -11 * pi / 12,
                                   5 * pi / 12, -5 * pi / 12, 7 * pi / 12, -7 * pi / 12 };
        const double inner[4] = { pi / 4, -pi / 4, 3 * pi / 4, -3 * pi / 4 };
        for (double a : outer)
            put(r2, a);
        for (double a : inner)
            put(r1, a);
        return true;
    }
    if (constellation == MOD_32APSK) {
        double g1, g2;
        switch (rate) {
        case C3_4: g1 = 2.84, g2 = 5.27; break;
        case C4_5: g1 = 2.72, g2 = 4.87; break;
        case C5_6: g1 = 2.64, g2 = 4.64; break;
        case C8_9: g1 = 2.54, g2 = 4.33; break;
        case C9_10: g1 = 2.53, g2 = 4.30; break;
        default: return false;
        }
        const double r1 = std::sqrt(8.0 / (1.0 + 3.0 * g1 * g1 + 4.0 * g2 * g2)), r2 = g1 * r1, r3 = g2 * r1;
        const double tab[32][2] = {
            { r2, pi / 4 }, { r2, 5 * pi / 12 }, { r2, -pi / 4 }, { r2, -5 * pi / 12 }, { r2, 3 * pi / 4 }, { r2, 7 * pi / 12 },
            { r2, -3 * pi / 4 }, { r2, -7 * pi / 12 }, { r3, pi / 8 }, { r3, 3 * pi / 8 }, { r3, -pi / 4 }, { r3, -pi / 2 },
            { r3, 3 * pi / 4 }, { r3, pi / 2 }, { r3, -7 * pi / 8 }, { r3, -5 * pi / 8 }, { r2, pi / 12 }, { r1, pi / 4 },
            { r2, -pi / 12 }, { r1, -pi / 4 }, { r2, 11 * pi / 12 }, { r1, 3 * pi / 4 }, { r2, -11 * pi / 12 }, { r1, -3 * pi / 4 },
            { r3, 0.0 }, { r3, pi / 4 }, { r3, -pi / 8 }, { r3, -3 * pi / 8 }, { r3, 7 * pi / 8 }, { r3, 5 * pi / 8 }, { r3, pi }, { r3, -3 * pi / 4 } };
        for (auto& t : tab)
            put(t[0], t[1]);
        return true;
    }
    return false;
}

namespace ldpc_cuda {
namespace {
dvbs2b200_code* g_code = nullptr; // a process-wide singleton, like the reference's ISA decoders
int g_simd = 32, g_n = 0;
std::vector<int32_t> g_ret;
} // namespace

int ldpc_dec_init(int standard, int framesize, int rate, int simd_size)
{
    ldpc_dec_shutdown();
    if (simd_size != 16 && simd_size != 32)
        return DVBS2B200_EINVAL;
    const char* env = getenv("DVBS2RX_B200_DEVICE");
    int rc = dvbs2b200_code_create(&g_code, env ? atoi(env) : 0, standard, framesize, rate);
    if (rc != DVBS2B200_OK)
        return rc;
    dvbs2b200_code_info info;
    dvbs2b200_code_info_get(g_code, &info);
    g_n = info.n_ldpc;
    g_simd = simd_size;
    g_ret.resize(simd_size);
    return DVBS2B200_OK;
}

int ldpc_dec_decode(void* /*buffer*/, int8_t* code, int trials)
{
    if (!g_code)
        return -1;
    // posteriors overwrite `code` in place; hard decisions are taken by the caller from them
    int rc = dvbs2b200_ldpc_decode(g_code, code, g_simd, trials, g_simd, /*OM_CODEWORD*/ 0, nullptr, code, g_ret.data());
    if (rc != DVBS2B200_OK)
        return -1;
    return g_ret[0];
}

void ldpc_dec_shutdown()
{
    if (g_code)
        dvbs2b200_code_destroy(g_code);
    g_code = nullptr;
}
} // namespace ldpc_cuda

// ---- C hooks so the blocks can be driven from the Python parity tests (ctypes) ------------------------
using namespace gr::dvbs2rx;
extern "C" {

struct blk_ldpc {
    ldpc_decoder_bb::sptr b;
    std::vector<int8_t> pdus;
    std::vector<uint64_t> pdu_frames;
};

void* blk_ldpc_make(int standard, int framesize, int rate, int outputmode, int max_trials, int batches, int want_pdu)
{
    try {
        blk_ldpc* h = new blk_ldpc();
        h->b = ldpc_decoder_bb::make((dvb_standard_t)standard, (dvb_framesize_t)framesize, (dvb_code_rate_t)rate,
                                     MOD_QPSK, (dvb_outputmode_t)outputmode, INFO_OFF, max_trials);
        h->b->set_batches_per_call(batches);
        if (want_pdu)
            h->b->set_llr_pdu_handler([h](const llr_pdu& p) {
                h->pdus.insert(h->pdus.end(), p.llr, p.llr + p.n_llr);
                h->pdu_frames.push_back(p.frame_cnt);
            });
        return h;
    } catch (...) {
        return nullptr;
    }
}
void blk_ldpc_free(void* h) { delete (blk_ldpc*)h; }
int blk_ldpc_output_multiple(void* h) { return ((blk_ldpc*)h)->b->output_multiple(); }
int blk_ldpc_forecast(void* h, int noutput_items)
{
    gr_vector_int req(1);
    ((blk_ldpc*)h)->b->forecast(noutput_items, req);
    return req[0];
}
int blk_ldpc_work(void* h, int noutput_items, const int8_t* in, int n_in, unsigned char* out, int* consumed)
{
    blk_ldpc* b = (blk_ldpc*)h;
    gr_vector_int ninput(1, n_in);
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    try {
        int r = b->b->general_work(noutput_items, ninput, ins, outs);
        *consumed = b->b->consumed();
        return r;
    } catch (...) {
        return -1000;
    }
}
unsigned blk_ldpc_average_trials(void* h) { return ((blk_ldpc*)h)->b->get_average_trials(); }
size_t blk_ldpc_pdu_bytes(void* h, int8_t* out, size_t cap)
{
    blk_ldpc* b = (blk_ldpc*)h;
    if (out && cap >= b->pdus.size())
        memcpy(out, b->pdus.data(), b->pdus.size());
    return b->pdus.size();
}

void* blk_bch_make(int standard, int framesize, int rate)
{
    try {
        auto* p = new bch_decoder_bb::sptr(
            bch_decoder_bb::make((dvb_standard_t)standard, (dvb_framesize_t)framesize, (dvb_code_rate_t)rate, OM_MESSAGE));
        return p;
    } catch (...) {
        return nullptr;
    }
}
void blk_bch_free(void* h) { delete (bch_decoder_bb::sptr*)h; }
int blk_bch_work(void* h, int noutput_items, const unsigned char* in, unsigned char* out, int* consumed)
{
    auto& b = *(bch_decoder_bb::sptr*)h;
    gr_vector_int ninput(1, 0);
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    try {
        int r = b->general_work(noutput_items, ninput, ins, outs);
        *consumed = b->consumed();
        return r;
    } catch (...) {
        return -1000;
    }
}
uint64_t blk_bch_frame_count(void* h) { return (*(bch_decoder_bb::sptr*)h)->get_frame_count(); }
uint64_t blk_bch_error_count(void* h) { return (*(bch_decoder_bb::sptr*)h)->get_error_count(); }

void* blk_bbdescrambler_make(int standard, int framesize, int rate)
{
    try {
        return new bbdescrambler_bb::sptr(bbdescrambler_bb::make((dvb_standard_t)standard, (dvb_framesize_t)framesize, (dvb_code_rate_t)rate));
    } catch (...) {
        return nullptr;
    }
}
void blk_bbdescrambler_free(void* h) { delete (bbdescrambler_bb::sptr*)h; }
int blk_bbdescrambler_work(void* h, int noutput_items, const unsigned char* in, unsigned char* out)
{
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    try {
        return (*(bbdescrambler_bb::sptr*)h)->work(noutput_items, ins, outs);
    } catch (...) {
        return -1000;
    }
}

void* blk_bbdeheader_make(int standard, int framesize, int rate)
{
    try {
        return new bbdeheader_bb::sptr(bbdeheader_bb::make((dvb_standard_t)standard, (dvb_framesize_t)framesize, (dvb_code_rate_t)rate));
    } catch (...) {
        return nullptr;
    }
}
void blk_bbdeheader_free(void* h) { delete (bbdeheader_bb::sptr*)h; }
int blk_bbdeheader_work(void* h, int noutput_items, int ninput_items, const unsigned char* in, unsigned char* out, int* consumed)
{
    auto& b = *(bbdeheader_bb::sptr*)h;
    gr_vector_int ninput(1, ninput_items);
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    try {
        int r = b->general_work(noutput_items, ninput, ins, outs);
        *consumed = b->consumed();
        return r;
    } catch (...) {
        return -1000;
    }
}
int blk_bbdeheader_forecast(void* h, int noutput_items)
{
    gr_vector_int req(1, 0);
    (*(bbdeheader_bb::sptr*)h)->forecast(noutput_items, req);
    return req[0];
}
void blk_bbdeheader_counters(void* h, uint64_t* out5)
{
    auto& b = *(bbdeheader_bb::sptr*)h;
    out5[0] = b->get_packet_count();
    out5[1] = b->get_error_count();
    out5[2] = b->get_bbframe_count();
    out5[3] = b->get_bbframe_drop_count();
    out5[4] = b->get_bbframe_gap_count();
}

void* blk_demap_make(int framesize, int rate, int constellation, char* err, int errcap)
{
    try {
        return new xfecframe_demapper_cb::sptr(
            xfecframe_demapper_cb::make((dvb_framesize_t)framesize, (dvb_code_rate_t)rate, (dvb_constellation_t)constellation));
    } catch (const std::exception& e) {
        if (err && errcap > 0) {
            strncpy(err, e.what(), errcap - 1);
            err[errcap - 1] = 0;
        }
        return nullptr;
    }
}
// the opt-in surface: 16APSK / 32APSK through the table-driven demapper, Es/N0 given explicitly
void* blk_demap_make_apsk(int framesize, int rate, int constellation, float es_n0_db, char* err, int errcap)
{
    try {
        auto* b = new xfecframe_demapper_cb::sptr(xfecframe_demapper_cb::make((dvb_framesize_t)framesize, (dvb_code_rate_t)rate,
                                                                             (dvb_constellation_t)constellation, /*apsk_opt_in=*/true));
        (*b)->set_es_n0_db(es_n0_db);
        return b;
    } catch (const std::exception& e) {
        if (err && errcap > 0) {
            strncpy(err, e.what(), errcap - 1);
            err[errcap - 1] = 0;
        }
        return nullptr;
    }
}
int blk_apsk_points(int constellation, int rate, float* points, int cap)
{
    std::vector<float> p;
    if (!apsk_points((dvb_constellation_t)constellation, (dvb_code_rate_t)rate, p) || (int)p.size() > cap)
        return -1;
    memcpy(points, p.data(), p.size() * sizeof(float));
    return (int)p.size() / 2;
}
void blk_demap_free(void* h) { delete (xfecframe_demapper_cb::sptr*)h; }
int blk_demap_work(void* h, int noutput_items, const float* in, int8_t* out, int* consumed)
{
    auto& b = *(xfecframe_demapper_cb::sptr*)h;
    gr_vector_int ninput(1, 0);
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    try {
        int r = b->general_work(noutput_items, ninput, ins, outs);
        *consumed = b->consumed();
        return r;
    } catch (...) {
        return -1000;
    }
}
float blk_demap_snr(void* h) { return (*(xfecframe_demapper_cb::sptr*)h)->get_snr(); }
void blk_demap_llr_pdu(void* h, long simd, uint64_t frame_cnt, const int8_t* llr, size_t n)
{
    llr_pdu p{ simd, frame_cnt, llr, n };
    (*(xfecframe_demapper_cb::sptr*)h)->handle_llr_pdu(p);
}

int blk_ldpc_cuda_init(int standard, int framesize, int rate, int simd) { return ldpc_cuda::ldpc_dec_init(standard, framesize, rate, simd); }
int blk_ldpc_cuda_decode(int8_t* code, int trials) { return ldpc_cuda::ldpc_dec_decode(nullptr, code, trials); }
void blk_ldpc_cuda_shutdown() { ldpc_cuda::ldpc_dec_shutdown(); }

} // extern "C"
