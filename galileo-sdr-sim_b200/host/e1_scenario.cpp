/* e1_scenario.cpp -- see e1_scenario.h.  Host C++ (no CUDA): RINEX 3 -> ephemerides -> per-block
 * pseudoranges -> e1_epoch_rec, plus the I/NAV page symbols the sample loop consumes.
 *
 * Written from the behavioural description of the reference (SURVEY.md sections 2, 3, 8a; line
 * citations below are to /root/reference).  Where the reference's behaviour hangs on evaluation
 * order of doubles, the expression is kept in that order; build with -O2 -ffp-contract=off.
 */
#include "e1_scenario.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../csrc/e1_core.h" /* e1_walk_up: exact count of code wraps inside a block */

namespace {

/* include/constants.h:53-66, 99-128, 168, 188 */
constexpr double kPi = 3.141592653589793;
constexpr double kRadToDeg = 57.2957795131;
constexpr double kLight = 2.99792458e8;
constexpr double kEarthRate = 7.2921151467e-5;
constexpr double kSqrtGM = 19964981.8432173887;
constexpr double kWgsA = 6378137.0, kWgsE = 0.0818191908426;
constexpr double kLambdaInit = 0.190293672798365;  /* LAMBDA_L1: carrier-phase initialisation (src/channel.cpp:98) */
constexpr double kLambdaE1 = 0.1902936727983649;   /* LAMBDA_E1: Doppler (src/gal-sig.cpp:318)                    */
constexpr double kCodeRate = 1.023e6, kCarrToCode = 0.0006493506493506494;
constexpr double kBlockDt = 0.10000002314200000;   /* src/galileo-sdr.cpp:347 */
constexpr int kMaxSat = 36;                        /* MAX_SAT */
constexpr int kSymbols = 500;

struct GalTime {
    int week;
    double sec;
};
struct CalDate {
    int y, m, d, hh, mm;
    double sec;
};

struct Ephemeris { /* the fields of ephem_t the path uses (include/structures.h:70-113) */
    int svid = 0, iode = 0, svhlth = 0;
    short flag = 0;
    GalTime toc{0, 0.0}, toe{0, 0.0};
    double af0 = 0, af1 = 0, af2 = 0, sqrta = 0, ecc = 0, inc0 = 0, omg0 = 0, aop = 0, m0 = 0, deltan = 0, omgdot = 0, idot = 0;
    double crc = 0, crs = 0, cuc = 0, cus = 0, cic = 0, cis = 0, bgde5a = 0, bgde5b = 0;
    double A = 0, n = 0, sq1e2 = 0, omgkdot = 0;
};

struct IonoUtc { /* ionoutc_t; vflg stays 0: the reference never sets it, so the
                    obliquity model is the one in force */
    int enable = 1;
    double A0 = 0, A1 = 0, ai0 = 0, ai1 = 0, ai2 = 0, ai3 = 0;
    int dtls = 0, tot = 0, wnt = 0, dtlsf = 0, dn = 0, wnlsf = 0;
};

struct Range {
    double range = 0, d = 0, azel[2] = {0, 0}, iono = 0;
};

/* ------------------------------------------------------------------ time (src/gnss-time.cpp) */
GalTime date_to_gal(const CalDate &t) /* :7-30 */
{
    static const int doy[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
    const int ye = t.y - 1980;
    int lpdays = ye / 4 + 1;
    if ((ye % 4) == 0 && t.m <= 2)
        lpdays--;
    const int de = ye * 365 + doy[t.m - 1] + t.d + lpdays - 6;
    GalTime g;
    g.week = de / 7;
    g.sec = (double)(de % 7) * 86400.0 + t.hh * 3600.0 + t.mm * 60.0 + t.sec;
    return g;
}

CalDate gal_to_date(const GalTime &g) /* :32-49 */
{
    CalDate t;
    const int c = (int)(7 * g.week + floor(g.sec / 86400.0) + 2444245.0) + 1537;
    const int d = (int)((c - 122.1) / 365.25);
    const int e = 365 * d + d / 4;
    const int f = (int)((c - e) / 30.6001);
    t.d = c - e - (int)(30.6001 * f);
    t.m = f - 1 - 12 * (f / 14);
    t.y = d - 4715 - ((7 + t.m) / 10);
    t.hh = ((int)(g.sec / 3600.0)) % 24;
    t.mm = ((int)(g.sec / 60.0)) % 60;
    t.sec = g.sec - 60.0 * floor(g.sec / 60.0);
    return t;
}

double gal_diff(const GalTime &a, const GalTime &b) /* subGalTime :80-88 */
{
    double dt = a.sec - b.sec;
    dt += (double)(a.week - b.week) * 604800.0;
    return dt;
}

/* ------------------------------------------------------------------ RINEX 3 (src/rinex.cpp) */
void d_to_e(char *s)
{
    for (; *s; s++)
        if (*s == 'D')
            *s = 'E';
}

/* one numeric field of a navigation record: present iff the line reaches past `probe` and that
   column is not blank (:83-96) */
double field(const char *line, size_t len, size_t probe, size_t start)
{
    if (len > probe && line[probe] != ' ')
        return strtod(line + start, nullptr);
    return 0.0;
}

bool read_rinex(const char *path, std::vector<Ephemeris> (&sv_eph)[kMaxSat], IonoUtc &iono)
{
    FILE *fp = fopen(path, "r");
    if (!fp)
        return false;
    char line[256];
    while (fgets(line, sizeof line, fp)) { /* header :118-158 */
        if (strncmp(line + 60, "END OF HEADER", 13) == 0)
            break;
        if (strncmp(line + 60, "IONOSPHERIC CORR", 16) == 0) {
            d_to_e(line);
            sscanf(line + 4, "%lf %lf %lf %lf", &iono.ai0, &iono.ai1, &iono.ai2, &iono.ai3);
        }
        if (strncmp(line + 60, "TIME SYSTEM CORR", 16) == 0 && strncmp(line, "GAUT", 4) == 0) {
            int d1 = 0, d2 = 0;
            d_to_e(line);
            const char keep = line[22];
            line[22] = 0;
            sscanf(line + 4, "%lf", &iono.A0);
            line[22] = keep;
            sscanf(line + 22, "%lf %d %d", &iono.A1, &d1, &d2);
            iono.tot = (unsigned char)(d1 >> 12);
            iono.wnt = (short)d2 >> 4;
            iono.wnlsf = (short)d2;
            iono.dtls = 18;
            iono.dtlsf = 18;
            iono.dn = 7;
        }
    }
    while (fgets(line, sizeof line, fp)) { /* records :161-236 */
        if (line[0] != 'E')
            continue;
        double v[39] = {0};
        CalDate when{};
        int isec = 0, svid = 0;
        d_to_e(line);
        size_t len = strlen(line);
        sscanf(line + 4, "%d %d %d %d %d %d", &when.y, &when.m, &when.d, &when.hh, &when.mm, &isec);
        when.sec = (double)isec;
        if (line[1] != ' ')
            sscanf(line + 1, "%2d", &svid);
        v[0] = field(line, len, 24, 23);
        v[1] = field(line, len, 43, 42);
        v[2] = field(line, len, 62, 61);
        for (int i = 0; i < 7; i++) {
            if (!fgets(line, sizeof line, fp))
                break;
            d_to_e(line);
            len = strlen(line);
            double *q = &v[i * 4 + 3];
            q[0] = field(line, len, 5, 4);
            q[1] = field(line, len, 24, 23);
            q[2] = field(line, len, 43, 42);
            q[3] = field(line, len, 62, 61);
        }
        Ephemeris e;
        e.svid = svid;
        e.toc = date_to_gal(when);
        e.af0 = v[0], e.af1 = v[1], e.af2 = v[2];
        e.sqrta = v[10], e.ecc = v[8], e.inc0 = v[15], e.omg0 = v[13], e.aop = v[17], e.m0 = v[6];
        e.deltan = v[5], e.omgdot = v[18], e.idot = v[19];
        e.crc = v[16], e.crs = v[4], e.cuc = v[7], e.cus = v[9], e.cic = v[12], e.cis = v[14];
        e.toe.sec = (int)(v[11] + 0.5);
        e.toe.week = (int)v[21];
        e.iode = (unsigned char)v[3];
        e.svhlth = (unsigned short)v[24];
        e.flag = (short)(unsigned short)v[20];
        if (e.flag != 517) /* I/NAV E1-B + E5b clock, :218 */
            continue;
        e.bgde5a = v[25];
        e.bgde5b = (e.flag & 0x2) ? v[25] : v[26];
        e.A = e.sqrta * e.sqrta;
        e.n = kSqrtGM / (e.sqrta * e.A) + e.deltan;
        e.sq1e2 = sqrt(1.0 - e.ecc * e.ecc);
        e.omgkdot = e.omgdot - kEarthRate;
        if (svid >= 1 && svid <= kMaxSat)
            sv_eph[svid - 1].push_back(e);
    }
    fclose(fp);
    return true;
}

/* first record whose clock epoch is within an hour of t (epoch_matcher, src/rinex.cpp:26-43) */
int match_epoch(const GalTime &t, const std::vector<Ephemeris> &v)
{
    for (size_t i = 0; i < v.size(); i++) {
        const double dt = gal_diff(t, v[i].toc);
        if (dt >= -3600.0 && dt < 3600.0)
            return (int)i;
    }
    return -1;
}

/* ------------------------------------------------------------------ geodesy */
double norm3(const double *x) { return sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); }

/* Geodetic latitude / longitude / height of an ECEF point by fixed-point iteration on the offset, along the polar
   axis, between the point and the centre of the prime-vertical circle through its footprint (offset = N e^2 sin(lat));
   stops when the offset moves by less than a millimetre.  The sequence of operations is the one of the reference's
   xyz2llh (src/geodesy.cpp:7-54): the records have to come out bit-identical (tests/test_host_scenario.py). */
void ecef_to_llh(const double *xyz, double *llh)
{
    const double ecc_sq = kWgsE * kWgsE, tolerance = 1.0e-3;
    if (norm3(xyz) < tolerance) { /* the Earth's centre: no direction to speak of */
        llh[0] = 0.0, llh[1] = 0.0, llh[2] = -kWgsA;
        return;
    }
    const double axis_dist_sq = xyz[0] * xyz[0] + xyz[1] * xyz[1];
    double offset = ecc_sq * xyz[2];
    double z_from_centre, range_from_centre, prime_vertical;
    for (;;) {
        z_from_centre = xyz[2] + offset;
        range_from_centre = sqrt(axis_dist_sq + z_from_centre * z_from_centre);
        const double sin_lat = z_from_centre / range_from_centre;
        prime_vertical = kWgsA / sqrt(1.0 - ecc_sq * sin_lat * sin_lat);
        const double next_offset = prime_vertical * ecc_sq * sin_lat;
        const bool settled = fabs(offset - next_offset) < tolerance;
        if (settled)
            break;
        offset = next_offset;
    }
    llh[0] = atan2(z_from_centre, sqrt(axis_dist_sq));
    llh[1] = atan2(xyz[1], xyz[0]);
    llh[2] = range_from_centre - prime_vertical;
}

void llh_to_ecef(const double *llh, double *xyz) /* :60-91 */
{
    const double a = kWgsA, e = kWgsE, e2 = e * e;
    const double clat = cos(llh[0]), slat = sin(llh[0]), clon = cos(llh[1]), slon = sin(llh[1]);
    const double d = e * slat;
    const double n = a / sqrt(1.0 - d * d);
    const double nph = n + llh[2];
    const double tmp = nph * clat;
    xyz[0] = tmp * clon;
    xyz[1] = tmp * slon;
    xyz[2] = ((1.0 - e2) * n + llh[2]) * slat;
}

void local_frame(const double *llh, double t[3][3]) /* ltcmat :97-118 */
{
    const double slat = sin(llh[0]), clat = cos(llh[0]), slon = sin(llh[1]), clon = cos(llh[1]);
    t[0][0] = -slat * clon, t[0][1] = -slat * slon, t[0][2] = clat;
    t[1][0] = -slon, t[1][1] = clon, t[1][2] = 0.0;
    t[2][0] = clat * clon, t[2][1] = clat * slon, t[2][2] = slat;
}

void az_el(const double *los, double t[3][3], double *azel) /* ecef2neu + neu2azel :125-152 */
{
    double neu[3];
    for (int i = 0; i < 3; i++)
        neu[i] = t[i][0] * los[0] + t[i][1] * los[1] + t[i][2] * los[2];
    azel[0] = atan2(neu[1], neu[0]);
    if (azel[0] < 0.0)
        azel[0] += (2.0 * kPi);
    const double ne = sqrt(neu[0] * neu[0] + neu[1] * neu[1]);
    azel[1] = atan2(neu[2], ne);
}

/* ---- broadcast orbit (Galileo OS SIS ICD 5.1.1, table 58: the user algorithm for the ephemeris; velocity by
 * differentiating it).  Split into the three steps of that table; every expression keeps the operand order of the
 * reference's satpos (src/geodesy.cpp:161-279), because f_carr and code_phase0 of the records are compared bit for bit
 * with the reference's trace and a last-bit difference in a satellite position shows there. */
double week_wrapped(double dt) /* time difference folded into half a week either side */
{
    if (dt > 302400.0)
        return dt - 604800.0;
    if (dt < -302400.0)
        return dt + 604800.0;
    return dt;
}

struct KeplerSolution {
    double sin_e, cos_e;    /* of the eccentric anomaly */
    double one_minus_ecos;  /* 1 - e cos E of the last iteration */
};

/* Newton's iteration on Kepler's equation from E = M, to 1e-14 rad */
KeplerSolution solve_kepler(double mean_anomaly, double ecc)
{
    KeplerSolution k;
    double anomaly = mean_anomaly, previous = anomaly + 1.0;
    k.one_minus_ecos = 0;
    for (int it = 0; fabs(anomaly - previous) > 1.0E-14 && it < 500; it++) {
        previous = anomaly;
        k.one_minus_ecos = 1.0 - ecc * cos(previous);
        anomaly = anomaly + (mean_anomaly - previous + ecc * sin(previous)) / k.one_minus_ecos;
    }
    k.sin_e = sin(anomaly);
    k.cos_e = cos(anomaly);
    return k;
}

struct InPlane { /* position in the orbital plane, the plane's orientation, and their rates */
    double x, y, x_rate, y_rate;
    double incl, incl_rate;
    double node;
};

InPlane in_plane_state(const Ephemeris &eph, double t_since_toe, const KeplerSolution &k)
{
    const double anomaly_rate = eph.n / k.one_minus_ecos;
    const double lat_arg0 = atan2(eph.sq1e2 * k.sin_e, k.cos_e - eph.ecc) + eph.aop;   /* true anomaly + argument of perigee */
    const double lat_arg0_rate = eph.sq1e2 * anomaly_rate / k.one_minus_ecos;
    const double s2 = sin(2.0 * lat_arg0), c2 = cos(2.0 * lat_arg0);                    /* second-harmonic corrections */
    const double lat_arg = lat_arg0 + eph.cus * s2 + eph.cuc * c2;
    const double sin_u = sin(lat_arg), cos_u = cos(lat_arg);
    const double lat_arg_rate = lat_arg0_rate * (1.0 + 2.0 * (eph.cus * c2 - eph.cuc * s2));
    const double radius = eph.A * k.one_minus_ecos + eph.crc * c2 + eph.crs * s2;
    const double radius_rate = eph.A * eph.ecc * k.sin_e * anomaly_rate + 2.0 * lat_arg0_rate * (eph.crs * c2 - eph.crc * s2);
    InPlane q;
    q.incl = eph.inc0 + eph.idot * t_since_toe + eph.cic * c2 + eph.cis * s2;
    q.incl_rate = eph.idot + 2.0 * lat_arg0_rate * (eph.cis * c2 - eph.cic * s2);
    q.x = radius * cos_u;
    q.y = radius * sin_u;
    q.x_rate = radius_rate * cos_u - q.y * lat_arg_rate;
    q.y_rate = radius_rate * sin_u + q.x * lat_arg_rate;
    q.node = eph.omg0 + t_since_toe * eph.omgkdot - kEarthRate * eph.toe.sec;           /* longitude of the node, Earth-fixed */
    return q;
}

/* position [m], velocity [m/s] in ECEF and clock bias / drift of a satellite at system time g */
void sat_state(const Ephemeris &eph, const GalTime &g, double *pos, double *vel, double *clk)
{
    const double t_since_toe = week_wrapped(g.sec - eph.toe.sec);
    const KeplerSolution k = solve_kepler(eph.m0 + eph.n * t_since_toe, eph.ecc);
    const InPlane q = in_plane_state(eph, t_since_toe, k);
    const double sin_i = sin(q.incl), cos_i = cos(q.incl), sin_node = sin(q.node), cos_node = cos(q.node);
    pos[0] = q.x * cos_node - q.y * cos_i * sin_node;
    pos[1] = q.x * sin_node + q.y * cos_i * cos_node;
    pos[2] = q.y * sin_i;
    const double y_rate_in_equator = q.y_rate * cos_i - q.y * sin_i * q.incl_rate;
    vel[0] = -eph.omgkdot * pos[1] + q.x_rate * cos_node - y_rate_in_equator * sin_node;
    vel[1] = eph.omgkdot * pos[0] + q.x_rate * sin_node + y_rate_in_equator * cos_node;
    vel[2] = q.y * cos_i * q.incl_rate + q.y_rate * sin_i;
    /* clock polynomial + relativistic term (F e sqrt(A) sin E) - the E1/E5b group delay, as the reference applies it */
    const double relativistic = -4.442807633E-10 * eph.ecc * eph.sqrta * k.sin_e;
    const double t_since_toc = week_wrapped(g.sec - eph.toc.sec);
    clk[0] = eph.af0 + t_since_toc * (eph.af1 + t_since_toc * eph.af2) + relativistic - eph.bgde5b;
    clk[1] = eph.af1 + 2.0 * t_since_toc * eph.af2;
}

/* obliquity-factor ionosphere (src/iono.cpp:9-19), the model in force with vflg == 0 (:37-40) */
double iono_delay(const IonoUtc &iono, const double *azel)
{
    if (!iono.enable)
        return 0.0;
    const double E = azel[1] / kPi;
    const double F = 1.0 + 16.0 * pow((0.53 - E), 3.0);
    return F * 5.0e-9 * kLight;
}

/* pseudorange of one satellite at receiver time g (computeRange, src/gal-sig.cpp:242-301) */
Range pseudorange(const Ephemeris &eph, const IonoUtc &iono, const GalTime &g, const double *xyz)
{
    double pos[3], vel[3], clk[2], los[3], llh[3], tmat[3][3];
    sat_state(eph, g, pos, vel, clk);
    for (int i = 0; i < 3; i++)
        los[i] = pos[i] - xyz[i];
    const double tau = norm3(los) / kLight;
    pos[0] -= vel[0] * tau;
    pos[1] -= vel[1] * tau;
    pos[2] -= vel[2] * tau;
    const double xrot = pos[0] + pos[1] * kEarthRate * tau;
    const double yrot = pos[1] - pos[0] * kEarthRate * tau;
    pos[0] = xrot;
    pos[1] = yrot;
    for (int i = 0; i < 3; i++)
        los[i] = pos[i] - xyz[i];
    Range r;
    const double range = norm3(los);
    r.d = range;
    r.range = range - kLight * clk[0];
    ecef_to_llh(xyz, llh);
    local_frame(llh, tmat);
    az_el(los, tmat, r.azel);
    r.iono = iono_delay(iono, r.azel);
    r.range += r.iono;
    return r;
}

bool visible(const Ephemeris &eph, const GalTime &g, const double *xyz, double mask_deg, double *azel) /* :319-343 */
{
    double llh[3], pos[3], vel[3], clk[2], los[3], tmat[3][3];
    ecef_to_llh(xyz, llh);
    local_frame(llh, tmat);
    sat_state(eph, g, pos, vel, clk);
    for (int i = 0; i < 3; i++)
        los[i] = pos[i] - xyz[i];
    az_el(los, tmat, azel);
    return azel[1] * kRadToDeg > mask_deg;
}

/* ------------------------------------------------------------------ I/NAV pages
 * (src/inav-msg.cpp, src/datatypes.cpp).  A page is built as a string of bits: 128-bit word +
 * reserved/SAR/spare, the odd half's header inserted at bit 114, CRC-24Q over the first 196 bits,
 * SSP; each 114-bit half (+6 tail zeros) is rate-1/2 K=7 encoded (second branch inverted),
 * 30x8 block-interleaved and prefixed with the 10-symbol sync pattern. */
struct BitWriter {
    int bits[256];
    int n = 0;
    BitWriter() { memset(bits, 0, sizeof bits); }
    /* encode_int_to_bits (datatypes.cpp:129-144): the value is masked with an `int` shift, which on
       x86-64 uses the count modulo 32; fields wider than 31 bits only ever carry zero */
    void put(long value, int width)
    {
        value &= (long)(int)((1u << (width & 31)) - 1u);
        for (int j = width - 1; j >= 0; j--)
            bits[n++] = (j < 32) ? (int)(((unsigned long)value >> j) & 1ul) : 0;
    }
    /* encode_double_to_bits (:146-159): the value goes through double -> int32 */
    void put32(double value, int width)
    {
        const int32_t q = (int32_t)value;
        for (int j = width - 1; j >= 0; j--)
            bits[n++] = (int)(((uint32_t)q >> j) & 1u);
    }
};

/* UnscaleULong / UnscaleLong / UnscaleInt / UnscaleUint (datatypes.cpp:32-71): round(|x| * 2^-scale)
   by mantissa arithmetic, sign applied afterwards, truncated to 32 bits */
unsigned long long unscale_mag(double value, int scale)
{
    unsigned long long b;
    memcpy(&b, &value, 8);
    const int exp = (int)((b >> 52) & 0x7ff);
    unsigned long long frac = b & 0xfffffffffffffULL;
    if (exp == 0 && frac == 0)
        return 0;
    frac |= 0x10000000000000ULL;
    const int shift = 1074 - exp + scale;
    frac += (1ULL << (shift & 63)); /* x86-64 shift semantics, as the reference build behaves */
    frac >>= ((shift + 1) & 63);
    return frac;
}
int unscale_int(double value, int scale)
{
    const long long mag = (long long)unscale_mag(value, scale);
    return (int)(signbit(value) ? -mag : mag);
}
unsigned int unscale_uint(double value, int scale) { return (unsigned int)unscale_mag(value, scale); }

/* CRC-24Q of a bit string (Galileo OS SIS ICD 5.1.9.3: generator 1864CFBh, zero initial value, bits in transmission
   order), one bit at a time.  The reference computes the same polynomial division bytewise with a table and a
   partial last byte (Crc24qEncode, src/inav-msg.cpp:134-162); that the two agree is held by the page symbols of the
   records (bit-identical to the reference's trace) and by the live-sky pages of tests/test_tv_pages.py. */
unsigned int crc24q_bits(const int *bits, int length)
{
    uint32_t rem = 0;
    for (int j = 0; j < length; j++) {
        rem ^= (uint32_t)(bits[j] & 1) << 23;
        rem <<= 1;
        if (rem & 0x1000000u)
            rem ^= 0x1864CFBu;
    }
    return rem & 0xFFFFFFu;
}

/* rate-1/2, K = 7 (G1 = 171o, G2 = 133o inverted), zero initial state; 120 bits -> 240 symbols
   (cnv_encd, inav-msg.cpp:58-118; taps include/constants.h:85-87) */
void conv_encode(const int *in, int n, int *out)
{
    static const int g1[7] = {1, 1, 1, 1, 0, 0, 1}, g2[7] = {1, 0, 1, 1, 0, 1, 1};
    int reg[7] = {0, 0, 0, 0, 0, 0, 0}; /* reg[0] newest */
    for (int t = 0; t < n; t++) {
        for (int j = 6; j > 0; j--)
            reg[j] = reg[j - 1];
        reg[0] = in[t];
        int a = 0, b = 0;
        for (int j = 0; j < 7; j++) {
            a ^= reg[j] & g1[j];
            b ^= reg[j] & g2[j];
        }
        out[2 * t] = a;
        out[2 * t + 1] = 1 - b;
    }
}

void half_page_symbols(const int *half120, int *out250) /* generateFrame, inav-msg.cpp:4-26 */
{
    static const int sync[10] = {0, 1, 0, 1, 1, 0, 0, 0, 0, 0};
    int fec[240];
    conv_encode(half120, 120, fec);
    for (int i = 0; i < 10; i++)
        out250[i] = sync[i];
    for (int r = 0; r < 8; r++)
        for (int c = 0; c < 30; c++)
            out250[10 + r * 30 + c] = fec[c * 8 + r];
}

/* generateINavMsg + generate_page (inav-msg.cpp:28-56, 165-409): the 500 symbols of the page that
   starts being sent at receiver time g */
void page_symbols(const GalTime &g, const Ephemeris &eph, const IonoUtc &iono, int *sym500)
{
    static const int word_of_slot[30] = {2, 4, 6, 7, 8, 17, 19, 16, 0, 0, 1, 3, 5, 0, 16,
                                         2, 4, 6, 9, 10, 17, 19, 16, 0, 0, 1, 3, 5, 0, 16}; /* galileo-sdr.h:32 */
    const int word = word_of_slot[((int)g.sec % 60) / 2];
    const int tow = (int)g.sec;
    BitWriter w;
    switch (word) {
    case 0:
        w.put(0, 8), w.put(2, 2), w.put(0, 88), w.put(g.week - 1024, 12), w.put(tow, 20);
        break;
    case 1:
        w.put(1, 8), w.put(eph.iode, 10), w.put((int)eph.toe.sec / 60, 14);
        w.put32((double)(long)unscale_int(eph.m0 / kPi, -31), 32);
        w.put32((double)unscale_uint(eph.ecc, -33), 32);
        w.put32((double)(long)unscale_int(eph.sqrta, -19), 32);
        w.put(0, 2);
        break;
    case 2:
        w.put(2, 8), w.put(eph.iode, 10);
        w.put32((double)(long)unscale_int(eph.omg0 / kPi, -31), 32);
        w.put32((double)(long)unscale_int(eph.inc0 / kPi, -31), 32);
        w.put32((double)(long)unscale_int(eph.aop / kPi, -31), 32);
        w.put32((double)(long)unscale_int(eph.idot / kPi, -43), 14);
        w.put(0, 2);
        break;
    case 3:
        w.put(3, 8), w.put(eph.iode, 10);
        w.put(unscale_int(eph.omgdot / kPi, -43), 24), w.put(unscale_int(eph.deltan / kPi, -43), 16);
        w.put(unscale_int(eph.cuc, -29), 16), w.put(unscale_int(eph.cus, -29), 16);
        w.put(unscale_int(eph.crc, -5), 16), w.put(unscale_int(eph.crs, -5), 16);
        w.put(32767, 8);
        break;
    case 4: {
        w.put(4, 8), w.put(eph.iode, 10), w.put(eph.svid, 6);
        w.put(unscale_int(eph.cic, -29), 16), w.put(unscale_int(eph.cis, -29), 16);
        const unsigned int toc_min = (unsigned int)(eph.toc.sec / 60);
        w.put((long)toc_min, 14);
        w.put(unscale_int(eph.af0, -34), 31), w.put(unscale_int(eph.af1, -46), 21), w.put(unscale_int(eph.af2, -59), 6);
        w.put(0, 2);
        break;
    }
    case 5:
        w.put(5, 8);
        w.put32((double)unscale_uint(iono.ai0, -2), 11);
        w.put32((double)(long)unscale_int(iono.ai1, -8), 11);
        w.put32((double)(long)unscale_int(iono.ai2, -15), 14);
        w.put(31, 5);
        w.put(unscale_int(eph.bgde5a, -32), 10), w.put(unscale_int(eph.bgde5b, -32), 10);
        w.put(eph.svhlth >> 7, 2), w.put(eph.svhlth >> 1, 2), w.put(eph.svhlth >> 5, 1), w.put(eph.svhlth, 1);
        w.put(g.week - 1024, 12), w.put(tow, 20), w.put(0, 23);
        break;
    case 6:
        w.put(6, 8);
        w.put32((double)(long)unscale_int(iono.A0, -30), 32);
        w.put32((double)(long)unscale_int(iono.A1, -50), 24);
        w.put(iono.dtls, 8);
        w.put((long)(iono.tot / 3600.0), 8);
        w.put(iono.wnt, 8), w.put(iono.wnlsf, 8), w.put(iono.dn, 3), w.put(iono.dtlsf, 8);
        w.put(tow, 20), w.put(0, 3);
        break;
    default: /* dummy word 63 */
        w.put(63, 8), w.put(0, 122);
        break;
    }
    w.put(0, 40), w.put(2796202, 22), w.put(0, 2);
    /* odd half's even/odd + page-type bits go in at bit 114 (shift_and_insert, :121-132) */
    for (int i = 239; i >= 116; i--)
        w.bits[i] = w.bits[i - 2];
    w.bits[114] = 1, w.bits[115] = 0;
    w.n += 2;
    w.put((long)crc24q_bits(w.bits, 196), 24);
    static const int ssp[3] = {4, 43, 47};
    w.put(ssp[word % 3], 8);
    int even[120] = {0}, odd[120] = {0};
    memcpy(even, w.bits, 114 * sizeof(int));
    memcpy(odd, w.bits + 114, 114 * sizeof(int));
    half_page_symbols(even, sym500);
    half_page_symbols(odd, sym500 + 250);
}

void pack_symbols(const int *sym500, uint8_t *out64)
{
    memset(out64, 0, E1_PAGE_BYTES);
    for (int k = 0; k < kSymbols; k++)
        if (sym500[k] > 0)
            out64[k >> 3] |= (uint8_t)(1u << (k & 7));
}

struct Channel { /* what galileo_task keeps per slot in channel_t */
    int prn = 0;
    int page[kSymbols];
    double rho0 = 0.0;      /* chan->rho0.range */
    double carr_phase0 = 0; /* value allocateChannel computed; goes out with the slot's first record */
    bool fresh = false;
};

} // namespace

struct e1h_scenario {
    e1h_options opt;
    std::vector<Ephemeris> eph[kMaxSat];
    IonoUtc iono;
    int current[kMaxSat];
    int slot_of_sv[kMaxSat];
    std::vector<Channel> chan;
    GalTime g0{0, 0.0}, grx{0, 0.0};
    double llh_deg[3];
    double xyz[3];
    double delt;
    int numd = 0, iumd = 1;
    std::vector<double> motion; /* optional: lat, lon [deg], height [m] per block index (e1h_set_motion) */
    double ant_pat[37];         /* receiver antenna pattern, linear, boresight angle 0:5:180 deg (src/galileo-sdr.cpp:50-54,365) */
    double elev_mask_deg = 10.0; /* allocateChannel's mask: the reference hard-codes 10 (src/channel.cpp:60) */

    /* allocateChannel (src/channel.cpp:21-122): runs on a COPY of the current-ephemeris indices */
    void allocate(const GalTime &g, const double *pos)
    {
        for (int sv = 0; sv < kMaxSat; sv++) {
            if (eph[sv].empty())
                continue;
            const int idx = match_epoch(g, eph[sv]);
            if (idx < 0)
                continue;
            const Ephemeris &e = eph[sv][idx];
            double azel[2];
            if (visible(e, g, pos, elev_mask_deg, azel)) {
                if (slot_of_sv[sv] != -1)
                    continue;
                int i;
                for (i = 0; i < (int)chan.size(); i++) {
                    Channel &c = chan[i];
                    if (c.prn != 0)
                        continue;
                    c.prn = sv + 1;
                    page_symbols(g, e, iono, c.page);
                    const Range r = pseudorange(e, iono, g, pos);
                    c.rho0 = r.range;
                    const double origin[3] = {0.0, 0.0, 0.0};
                    const Range ref = pseudorange(e, iono, g, origin);
                    const double phase_ini = (2.0 * ref.range - r.range) / kLambdaInit;
                    c.carr_phase0 = phase_ini - floor(phase_ini);
                    c.fresh = true;
                    if (opt.verbose)
                        fprintf(stderr, "%02d %6.1f %5.1f %11.1f %5.5f\n", c.prn, azel[0] * kRadToDeg, azel[1] * kRadToDeg, c.rho0, g.sec);
                    break;
                }
                if (i < (int)chan.size())
                    slot_of_sv[sv] = i;
            } else if (slot_of_sv[sv] >= 0) {
                chan[slot_of_sv[sv]].prn = 0;
                slot_of_sv[sv] = -1;
            }
        }
    }
};

extern "C" {

void e1h_default_options(e1h_options *o)
{
    memset(o, 0, sizeof *o);
    o->llh[0] = 42.3601, o->llh[1] = -71.0589, o->llh[2] = 2; /* src/main.cpp:189-191 */
    o->iduration = 3000;                                       /* USER_MOTION_SIZE */
    o->iono_enable = 1;
    o->max_chan = 16;
    o->fs_hz = (double)2.6e6f;
    o->samples_per_epoch = 260000;
}

e1h_scenario *e1h_open(const e1h_options *o, char *err, int err_len)
{
    auto fail = [&](const std::string &m) -> e1h_scenario * {
        if (err && err_len > 0)
            snprintf(err, (size_t)err_len, "%s", m.c_str());
        return nullptr;
    };
    if (!o || o->max_chan < 1 || o->max_chan > E1B200_MAX_CHAN || o->iduration < 1)
        return fail("bad options");
    e1h_scenario *s = new e1h_scenario();
    {
        /* attenuation in dB per 5 degrees off boresight: the pattern table gps-sdr-sim and this reference ship
           (src/galileo-sdr.cpp:50-54), turned into amplitude factors as at :365 */
        static const double ant_pat_db[37] = {0.00, 0.00, 0.22, 0.44, 0.67, 1.11, 1.56, 2.00, 2.44, 2.89, 3.56, 4.22, 4.89,
                                              5.56, 6.22, 6.89, 7.56, 8.22, 8.89, 9.78, 10.67, 11.56, 12.44, 13.33, 14.44, 15.56,
                                              16.67, 17.78, 18.89, 20.00, 21.33, 22.67, 24.00, 25.56, 27.33, 29.33, 31.56};
        for (int i = 0; i < 37; i++)
            s->ant_pat[i] = pow(10.0, -ant_pat_db[i] / 20.0);
        if (o->elev_mask_deg != 0.0)
            s->elev_mask_deg = o->elev_mask_deg;
    }
    s->opt = *o;
    s->iono.enable = o->iono_enable;
    if (!read_rinex(o->navfile, s->eph, s->iono)) {
        delete s;
        return fail(std::string("Error opening file: ") + o->navfile);
    }
    /* earliest / latest usable clock epochs (src/galileo-sdr.cpp:230-273) */
    GalTime gmin{0, 0.0}, gmax{0, 0.0};
    bool any = false;
    for (int sv = 0; sv < kMaxSat && !any; sv++)
        if (!s->eph[sv].empty()) {
            gmin = s->eph[sv][0].toc;
            any = true;
        }
    if (!any) {
        delete s;
        return fail("no Galileo I/NAV (E1-B/E5b, data source 517) records in the navigation file");
    }
    for (int sv = 0; sv < kMaxSat; sv++) {
        const size_t n = s->eph[sv].size();
        if (n >= 2 && s->eph[sv][n - 2].toc.sec > gmax.sec)
            gmax = s->eph[sv][n - 2].toc;
    }
    if (o->have_start) { /* -t: src/main.cpp:261-272, src/gnss-time.cpp:141-157 */
        CalDate t0{o->y, o->m, o->d, o->hh, o->mm, floor(o->sec)};
        if (t0.y <= 1980 || t0.m < 1 || t0.m > 12 || t0.d < 1 || t0.d > 31 || t0.hh < 0 || t0.hh > 23 || t0.mm < 0 || t0.mm > 59 ||
            o->sec < 0.0 || o->sec >= 60.0) {
            delete s;
            return fail("ERROR: Invalid date and time.");
        }
        s->g0 = date_to_gal(t0);
        if (gal_diff(s->g0, gmin) < 0.0 || gal_diff(gmax, s->g0) < 0.0) {
            delete s;
            return fail("ERROR: Invalid start time.");
        }
    } else {
        s->g0 = gmin;
    }
    s->chan.resize((size_t)o->max_chan);
    for (int sv = 0; sv < kMaxSat; sv++)
        s->slot_of_sv[sv] = -1;
    s->delt = 1.0 / o->fs_hz;
    s->numd = o->iduration;
    /* position: degrees -> radians with the reference's truncated constant (src/galileo-sdr.cpp:211-213) */
    for (int i = 0; i < 3; i++)
        s->llh_deg[i] = o->llh[i];
    double llh[3] = {o->llh[0] / kRadToDeg, o->llh[1] / kRadToDeg, o->llh[2]};
    llh_to_ecef(llh, s->xyz);
    s->grx = s->g0;
    for (int sv = 0; sv < kMaxSat; sv++)
        s->current[sv] = match_epoch(s->grx, s->eph[sv]); /* :304-306 */
    if (o->verbose) {
        const CalDate tl = gal_to_date(s->g0);
        fprintf(stderr, "xyz = %11.1f, %11.1f, %11.1f\n", s->xyz[0], s->xyz[1], s->xyz[2]);
        fprintf(stderr, "llh = %11.6f, %11.6f, %11.1f\n", llh[0] * kRadToDeg, llh[1] * kRadToDeg, llh[2]);
        fprintf(stderr, "Duration = %.1f [sec]\n", ((double)s->numd) / 10.0);
        fprintf(stderr, "Start = %4d/%02d/%02d,%02d:%02d:%02.0f (%d:%.0f)\n", tl.y, tl.m, tl.d, tl.hh, tl.mm, tl.sec, s->g0.week, s->g0.sec);
    }
    s->grx.sec = s->grx.sec + kBlockDt; /* :347-352 */
    s->allocate(s->grx, s->xyz);
    s->grx.sec = s->grx.sec + kBlockDt; /* :436 */
    s->iumd = 1;
    return s;
}

void e1h_close(e1h_scenario *s) { delete s; }

int e1h_total_epochs(const e1h_scenario *s) { return s ? s->numd - 1 : 0; }

int e1h_set_location(e1h_scenario *s, double lat_deg, double lon_deg, double height_m)
{
    if (!s)
        return -1;
    s->llh_deg[0] = lat_deg, s->llh_deg[1] = lon_deg, s->llh_deg[2] = height_m;
    return 0;
}

int e1h_set_motion(e1h_scenario *s, int n_blocks, const double *llh_deg)
{
    if (!s || n_blocks < 0 || (n_blocks && !llh_deg))
        return -1;
    s->motion.assign(llh_deg, llh_deg + (size_t)n_blocks * 3);
    return 0;
}

int e1h_ecef_to_llh_deg(const double *xyz, double *llh_deg)
{
    double llh[3];
    ecef_to_llh(xyz, llh);
    llh_deg[0] = llh[0] * kRadToDeg, llh_deg[1] = llh[1] * kRadToDeg, llh_deg[2] = llh[2];
    return 0;
}

int e1h_next(e1h_scenario *s, int n, e1_epoch_rec *recs, double *grx_sec) { return e1h_next_ex(s, n, recs, nullptr, grx_sec); }

int e1h_next_ex(e1h_scenario *s, int n, e1_epoch_rec *recs, e1_range_rec *ranges, double *grx_sec)
{
    if (!s || (!recs && !ranges) || n < 0)
        return -1;
    const int max_chan = s->opt.max_chan, n_samp = s->opt.samples_per_epoch;
    std::vector<e1_epoch_rec> scratch;
    if (!recs)
        scratch.resize((size_t)max_chan);
    int done = 0;
    for (; done < n && s->iumd < s->numd; done++, s->iumd++) {
        e1_epoch_rec *out = recs ? recs + (size_t)done * max_chan : scratch.data();
        e1_range_rec *rout = ranges ? ranges + (size_t)done * max_chan : nullptr;
        memset(out, 0, sizeof(e1_epoch_rec) * (size_t)max_chan);
        if (rout)
            memset(rout, 0, sizeof(e1_range_rec) * (size_t)max_chan);
        /* position of this block: the location thread's degrees (llhr, include/socket.h:69,165-178 --
           here e1h_set_location or the motion table), converted again (src/galileo-sdr.cpp:443-448) */
        if ((size_t)s->iumd * 3 + 2 < s->motion.size())
            for (int i = 0; i < 3; i++)
                s->llh_deg[i] = s->motion[(size_t)s->iumd * 3 + i];
        double llh[3] = {s->llh_deg[0] / kRadToDeg, s->llh_deg[1] / kRadToDeg, s->llh_deg[2]};
        llh_to_ecef(llh, s->xyz);
        if (grx_sec)
            grx_sec[done] = s->grx.sec;
        for (int i = 0; i < max_chan; i++) {
            Channel &c = s->chan[i];
            if (c.prn <= 0)
                continue;
            const int sv = c.prn - 1;
            const Ephemeris &eph = s->eph[sv][s->current[sv] < 0 ? 0 : s->current[sv]];
            const Range rho = pseudorange(eph, s->iono, s->grx, s->xyz);
            /* computeCodePhase (src/gal-sig.cpp:308-347) */
            const double rhorate = (rho.range - c.rho0) / kBlockDt;
            const double f_carr = (-rhorate / kLambdaE1);
            const double f_code = kCodeRate + f_carr * kCarrToCode;
            double ms = (s->grx.sec - rho.range / kLight) * 1000.0;
            const int ipage = ms / 2000.0;
            ms -= ipage * 2000;
            int ibit = (unsigned int)ms / 4;
            ms -= ibit * 4;
            const double code_phase = ms / 4 * E1_CODE_LEN;
            ibit = (ibit + (E1_SYM_PER_PAGE / 2)) % E1_SYM_PER_PAGE;
            if (rout) { /* the same block as pseudoranges: computeCodePhase is then evaluated on the device */
                rout[i].prn = c.prn;
                rout[i].rho_prev = c.rho0;
                rout[i].rho_cur = rho.range;
                rout[i].grx_sec = s->grx.sec;
            }
            c.rho0 = rho.range;

            e1_epoch_rec &r = out[i];
            r.prn = c.prn;
            /* gain[i] (src/galileo-sdr.cpp:469-477): free-space loss relative to 20 200 km times the receiver
               antenna pattern at the boresight angle, scaled by 2^7.  The reference computes it every block and
               never applies it (:520-521); the record carries it for E1B200_CFG_GAIN. */
            {
                const double path_loss = 20200000.0 / rho.d;
                const int ibs = (int)((90.0 - rho.azel[1] * kRadToDeg) / 5.0);
                r.gain_q7 = (int)(path_loss * s->ant_pat[ibs < 0 ? 0 : (ibs > 36 ? 36 : ibs)] * 128.0);
            }
            r.ibit0 = ibit;
            r.code_phase0 = code_phase;
            r.f_code = f_code;
            r.f_carr = f_carr;
            if (c.fresh) {
                r.flags = E1_REC_SET_PHASE;
                r.carr_phase_init = c.carr_phase0;
                c.fresh = false;
            }
            if (rout) {
                rout[i].flags = r.flags | ((uint32_t)r.gain_q7 << 8); /* e1_range_rec: gain in the upper flag bits */
                rout[i].carr_phase_init = r.carr_phase_init;
            }
            pack_symbols(c.page, r.page_cur);
            /* does the symbol counter pass 499 inside this block (src/galileo-sdr.cpp:491-506)?  Count the
               code wraps the sample loop will see, exactly (same roundings, e1_core.h walker). */
            int wraps = 0;
            {
                const double sc = f_code * s->delt;
                double cp = code_phase;
                if (cp >= (double)E1_CODE_LEN) {
                    cp -= (double)E1_CODE_LEN;
                    wraps++;
                }
                int64_t k = 0;
                while (k < n_samp) {
                    int w = 0;
                    cp = e1_walk_up(cp, sc, (double)E1_CODE_LEN, &k, n_samp, &w);
                    if (w && k < n_samp) /* a wrap "at sample n_samp" is never tested by the loop */
                        wraps++;
                }
            }
            if (ibit + wraps >= E1_SYM_PER_PAGE) {
                int next[kSymbols];
                page_symbols(s->grx, eph, s->iono, next);
                pack_symbols(next, r.page_next);
                memcpy(c.page, next, sizeof next);
            } else {
                memcpy(r.page_next, r.page_cur, E1_PAGE_BYTES);
            }
            if (rout) {
                memcpy(rout[i].page_cur, r.page_cur, E1_PAGE_BYTES);
                memcpy(rout[i].page_next, r.page_next, E1_PAGE_BYTES);
            }
        }
        /* every 30 s of receiver time: re-match ephemerides, add / drop satellites (:545-562) */
        const int igrx = (int)(s->grx.sec * 10.0 + 0.5);
        if ((int)fmodf((float)igrx, 300) == 0) {
            for (int sv = 0; sv < kMaxSat; sv++)
                s->current[sv] = match_epoch(s->grx, s->eph[sv]);
            s->allocate(s->grx, s->xyz);
        }
        s->grx.sec = s->grx.sec + kBlockDt;
    }
    return done;
}

int e1h_page_symbols(const e1h_scenario *s, int prn, double grx_sec, int week, int *symbols500)
{
    if (!s || prn < 1 || prn > kMaxSat || s->eph[prn - 1].empty())
        return -1;
    GalTime g{week, grx_sec};
    const int idx = match_epoch(g, s->eph[prn - 1]);
    if (idx < 0)
        return -1;
    page_symbols(g, s->eph[prn - 1][idx], s->iono, symbols500);
    return 0;
}

unsigned int e1h_crc24q_bits(const int *bits, int length) { return crc24q_bits(bits, length); }

/* The channel coding of a page on its own (what page_symbols does after the word is laid out): two half pages of
   114 bits -> 6 tail zeros, rate-1/2 code, 30 x 8 interleaver, sync pattern -> 500 symbols.  For the tests that push
   live-sky pages (the reference's tv/ vectors) through this encoder and the receiver stand-in's decoder. */
int e1h_encode_page(const int *even114, const int *odd114, int *symbols500)
{
    if (!even114 || !odd114 || !symbols500)
        return -1;
    int even[120] = {0}, odd[120] = {0};
    memcpy(even, even114, 114 * sizeof(int));
    memcpy(odd, odd114, 114 * sizeof(int));
    half_page_symbols(even, symbols500);
    half_page_symbols(odd, symbols500 + 250);
    return 0;
}

} /* extern "C" */
