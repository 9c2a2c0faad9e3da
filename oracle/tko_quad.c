/* tko_quad.c -- oracle restatement of the quadrature layer.
 *
 * TEST INFRASTRUCTURE ONLY (see tko_fp.h).
 * Follows f90src/helpers/integration.f90:
 *   integrate :18-49, adaptive_integrate :51-64, qk21 :77-224, iqags :231-244,
 *   qags :246-611, qsort :616-729, qextr :732-904   (QUADPACK, control flow kept 1:1,
 *   including the float32 literal 1.0e-05 at :464 and the ignored ifail).
 * Arrays are 1-based here too (index 0 unused) so the index arithmetic reads like
 * the reference.
 */
#include "tko.h"
#include "tko_fp.h"

#define LIMIT 500

static const double xgk[12] = { 0.0,
    9.956571630258081e-01, 9.739065285171717e-01, 9.301574913557082e-01, 8.650633666889845e-01,
    7.808177265864169e-01, 6.794095682990244e-01, 5.627571346686047e-01, 4.333953941292472e-01,
    2.943928627014602e-01, 1.488743389816312e-01, 0.000000000000000 };
static const double wgk[12] = { 0.0,
    1.169463886737187e-02, 3.255816230796473e-02, 5.475589657435200e-02, 7.503967481091995e-02,
    9.312545458369761e-02, 1.093871588022976e-01, 1.234919762620659e-01, 1.347092173114733e-01,
    1.427759385770601e-01, 1.477391049013385e-01, 1.494455540029169e-01 };
static const double wg[6] = { 0.0,
    6.667134430868814e-02, 1.494513491505806e-01, 2.190863625159820e-01, 2.692667193099964e-01,
    2.955242247147529e-01 };

/* integration.f90:77-224 */
void tko_qk21(tko_f1 f, void *ctx, double a, double b,
              double *result, double *abserr, double *resabs_out, double *resasc_out)
{
    double fv1[11], fv2[11];
    double centr = 5.0e-01 * (a + b);
    double hlgth = 5.0e-01 * (b - a);
    double dhlgth = fabs(hlgth);
    double resg = 0.0;
    double fc = f(centr, ctx);
    double resk = wgk[11] * fc;
    double resabs = fabs(resk);
    for (int j = 1; j <= 5; j++) {
        int jtw = 2 * j;
        double absc = hlgth * xgk[jtw];
        double fval1 = f(centr - absc, ctx);
        double fval2 = f(centr + absc, ctx);
        fv1[jtw] = fval1;
        fv2[jtw] = fval2;
        double fsum = fval1 + fval2;
        resg = resg + wg[j] * fsum;
        resk = resk + wgk[jtw] * fsum;
        resabs = resabs + wgk[jtw] * (fabs(fval1) + fabs(fval2));
    }
    for (int j = 1; j <= 5; j++) {
        int jtwm1 = 2 * j - 1;
        double absc = hlgth * xgk[jtwm1];
        double fval1 = f(centr - absc, ctx);
        double fval2 = f(centr + absc, ctx);
        fv1[jtwm1] = fval1;
        fv2[jtwm1] = fval2;
        double fsum = fval1 + fval2;
        resk = resk + wgk[jtwm1] * fsum;
        resabs = resabs + wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
    }
    double reskh = resk * 5.0e-01;
    double resasc = wgk[11] * fabs(fc - reskh);
    for (int j = 1; j <= 10; j++)
        resasc = resasc + wgk[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    *result = resk * hlgth;
    resabs = resabs * dhlgth;
    resasc = resasc * dhlgth;
    double err = fabs((resk - resg) * hlgth);
    if (resasc != 0.0 && err != 0.0)
        err = resasc * tko_min(1.0, tko_pow15_capped(2.0e+02 * err / resasc));
    if (resabs > TKO_TINY / (5.0e+01 * TKO_EPS))
        err = tko_max((TKO_EPS * 5.0e+01) * resabs, err);
    *abserr = err;
    *resabs_out = resabs;
    *resasc_out = resasc;
}

/* integration.f90:616-729 */
static void qsort_(int limit, int last, int *maxerr, double *ermax, const double *elist, int *iord, int *nrmax)
{
    int i, ibeg, isucc, j, jbnd, jupbn, k;
    double errmax, errmin;
    if (last <= 2) {
        iord[1] = 1;
        iord[2] = 2;
        goto L90;
    }
    errmax = elist[*maxerr];
    {
        int ntrip = *nrmax - 1;                 /* Fortran fixes the trip count on entry */
        for (i = 1; i <= ntrip; i++) {
            isucc = iord[*nrmax - 1];
            if (errmax <= elist[isucc]) break;
            iord[*nrmax] = isucc;
            *nrmax = *nrmax - 1;
        }
    }
    jupbn = last;
    if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
    errmin = elist[last];
    jbnd = jupbn - 1;
    ibeg = *nrmax + 1;
    for (i = ibeg; i <= jbnd; i++) {
        isucc = iord[i];
        if (errmax >= elist[isucc]) goto L60;
        iord[i - 1] = isucc;
    }
    iord[jbnd] = *maxerr;
    iord[jupbn] = last;
    goto L90;
L60:
    iord[i - 1] = *maxerr;
    k = jbnd;
    for (j = i; j <= jbnd; j++) {
        isucc = iord[k];
        if (errmin < elist[isucc]) goto L80;
        iord[k + 1] = isucc;
        k = k - 1;
    }
    iord[i] = last;
    goto L90;
L80:
    iord[k + 1] = last;
L90:
    *maxerr = iord[*nrmax];
    *ermax = elist[*maxerr];
}

/* integration.f90:732-904 */
static void qextr_(int *n, double *epstab, double *result, double *abserr, double *res3la, int *nres)
{
    int i, ib, ib2, ie, indx, k1, k2, k3, limexp, newelm, num;
    double delta1, delta2, delta3, epsinf, error, err1, err2, err3, e0, e1, e1abs, e2, e3, res, ss,
           tol1, tol2, tol3;
    *nres = *nres + 1;
    *abserr = TKO_HUGE;
    *result = epstab[*n];
    if (*n < 3) goto L100;
    limexp = 50;
    epstab[*n + 2] = epstab[*n];
    newelm = (*n - 1) / 2;
    epstab[*n] = TKO_HUGE;
    num = *n;
    k1 = *n;
    for (i = 1; i <= newelm; i++) {
        k2 = k1 - 1;
        k3 = k1 - 2;
        res = epstab[k1 + 2];
        e0 = epstab[k3];
        e1 = epstab[k2];
        e2 = res;
        e1abs = fabs(e1);
        delta2 = e2 - e1;
        err2 = fabs(delta2);
        tol2 = tko_max(fabs(e2), e1abs) * TKO_EPS;
        delta3 = e1 - e0;
        err3 = fabs(delta3);
        tol3 = tko_max(e1abs, fabs(e0)) * TKO_EPS;
        if (err2 <= tol2 && err3 <= tol3) {
            *result = res;
            *abserr = err2 + err3;
            goto L100;
        }
        e3 = epstab[k1];
        epstab[k1] = e1;
        delta1 = e1 - e3;
        err1 = fabs(delta1);
        tol1 = tko_max(e1abs, fabs(e3)) * TKO_EPS;
        if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) goto L20;
        ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
        epsinf = fabs(ss * e1);
        if (epsinf > 1.0e-04) goto L30;
L20:
        *n = i + i - 1;
        break;
L30:
        res = e1 + 1.0 / ss;
        epstab[k1] = res;
        k1 = k1 - 2;
        error = err2 + fabs(res - e2) + err3;
        if (error <= *abserr) {
            *abserr = error;
            *result = res;
        }
    }
    if (*n == limexp) *n = 2 * (limexp / 2) - 1;
    if ((num / 2) * 2 == num) ib = 2; else ib = 1;
    ie = newelm + 1;
    for (i = 1; i <= ie; i++) {
        ib2 = ib + 2;
        epstab[ib] = epstab[ib2];
        ib = ib2;
    }
    if (num != *n) {
        indx = num - *n + 1;
        for (i = 1; i <= *n; i++) {
            epstab[i] = epstab[indx];
            indx = indx + 1;
        }
    }
    if (*nres < 4) {
        res3la[*nres] = *result;
        *abserr = TKO_HUGE;
    } else {
        *abserr = fabs(*result - res3la[3]) + fabs(*result - res3la[2]) + fabs(*result - res3la[1]);
        res3la[1] = res3la[2];
        res3la[2] = res3la[3];
        res3la[3] = *result;
    }
L100:
    *abserr = tko_max(*abserr, 0.5 * TKO_EPS * fabs(*result));
}

/* integration.f90:246-611 */
void tko_qags(tko_f1 f, void *ctx, double a, double b, double epsabs, double epsrel,
              double *result_out, double *abserr_out, int *neval_out, int *ifail_out, int *last_out)
{
    const int limit = LIMIT;
    int extrap, noext;
    int id, ierro, iord[LIMIT + 1], iroff1, iroff2, iroff3, jupbnd, k, ksgn, ktmin,
        last, maxerr, nres, nrmax, numrl2, ifail;
    double abseps, alist[LIMIT + 1], area, area1, area12, area2, a1, a2, blist[LIMIT + 1], b1, b2,
           correc, defabs, defab1, defab2, dres, elist[LIMIT + 1], erlarg, erlast, errbnd,
           errmax, error1, error2, erro12, errsum, ertest, resabs, reseps, res3la[4],
           rlist[LIMIT + 1], rlist2[53], small, result, abserr;

    ifail = 0;
    last = 0;
    result = 0.0;
    abserr = 0.0;
    alist[1] = a;
    blist[1] = b;
    rlist[1] = 0.0;
    elist[1] = 0.0;
    if (epsabs < 0.0 && epsrel < 0.0) {
        ifail = 6;
        /* the reference returns here without touching neval/ifail outputs */
        *result_out = result; *abserr_out = abserr;
        if (neval_out) *neval_out = 0;
        if (ifail_out) *ifail_out = ifail;
        if (last_out) *last_out = last;
        return;
    }

    ierro = 0;
    tko_qk21(f, ctx, a, b, &result, &abserr, &defabs, &resabs);

    dres = fabs(result);
    errbnd = tko_max(epsabs, epsrel * dres);
    last = 1;
    rlist[1] = result;
    elist[1] = abserr;
    iord[1] = 1;

    if (abserr <= 1.0e+02 * TKO_EPS * defabs && abserr > errbnd) ifail = 2;
    if (limit == 1) ifail = 1;
    if (ifail != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) goto L140;

    rlist2[1] = result;
    errmax = abserr;
    maxerr = 1;
    area = result;
    errsum = abserr;
    abserr = TKO_HUGE;
    nrmax = 1;
    nres = 0;
    numrl2 = 2;
    ktmin = 0;
    extrap = 0;
    noext = 0;
    iroff1 = 0;
    iroff2 = 0;
    iroff3 = 0;
    correc = 0;
    erlarg = 0;
    small = 0;
    ertest = 0;

    if (dres >= (1.0 - 5.0e+01 * TKO_EPS) * defabs) ksgn = 1; else ksgn = -1;

    for (last = 2; last <= limit; last++) {
        a1 = alist[maxerr];
        b1 = 5.0e-01 * (alist[maxerr] + blist[maxerr]);
        a2 = b1;
        b2 = blist[maxerr];
        erlast = errmax;
        tko_qk21(f, ctx, a1, b1, &area1, &error1, &resabs, &defab1);
        tko_qk21(f, ctx, a2, b2, &area2, &error2, &resabs, &defab2);

        area12 = area1 + area2;
        erro12 = error1 + error2;
        errsum = errsum + erro12 - errmax;
        area = area + area12 - rlist[maxerr];

        if (defab1 == error1 || defab2 == error2) goto L15;
        /* `1.0e-05` is a default-kind (float32) literal at integration.f90:464 */
        if (fabs(rlist[maxerr] - area12) > (double)1.0e-05f * fabs(area12) || erro12 < 9.9e-01 * errmax) goto L10;
        if (extrap) iroff2 = iroff2 + 1; else iroff1 = iroff1 + 1;
L10:
        if (last > 10 && erro12 > errmax) iroff3 = iroff3 + 1;
L15:
        rlist[maxerr] = area1;
        rlist[last] = area2;
        errbnd = tko_max(epsabs, epsrel * fabs(area));

        if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ifail = 2;
        if (iroff2 >= 5) ierro = 3;
        if (last == limit) ifail = 1;
        if (tko_max(fabs(a1), fabs(b2)) <= (1.0 + 1.0e+03 * TKO_EPS) * (fabs(a2) + 1.0e+03 * TKO_TINY)) ifail = 4;

        if (error2 <= error1) {
            alist[last] = a2;
            blist[maxerr] = b1;
            blist[last] = b2;
            elist[maxerr] = error1;
            elist[last] = error2;
        } else {
            alist[maxerr] = a2;
            alist[last] = a1;
            blist[last] = b1;
            rlist[maxerr] = area2;
            rlist[last] = area1;
            elist[maxerr] = error2;
            elist[last] = error1;
        }

        qsort_(limit, last, &maxerr, &errmax, elist, iord, &nrmax);
        if (errsum <= errbnd) goto L115;
        if (ifail != 0) break;
        if (last == 2) goto L80;
        if (noext) goto L90;
        erlarg = erlarg - erlast;
        if (fabs(b1 - a1) > small) erlarg = erlarg + erro12;
        if (extrap) goto L40;

        if (fabs(blist[maxerr] - alist[maxerr]) > small) goto L90;
        extrap = 1;
        nrmax = 2;
L40:
        if (ierro != 3 && erlarg > ertest) {
            id = nrmax;
            jupbnd = last;
            if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
            for (k = id; k <= jupbnd; k++) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) goto L90;
                nrmax = nrmax + 1;
            }
        }
        /* perform extrapolation (label 60) */
        numrl2 = numrl2 + 1;
        rlist2[numrl2] = area;
        qextr_(&numrl2, rlist2, &reseps, &abseps, res3la, &nres);
        ktmin = ktmin + 1;
        if (ktmin > 5 && abserr < 1.0e-03 * errsum) ifail = 5;
        if (abseps < abserr) {
            ktmin = 0;
            abserr = abseps;
            result = reseps;
            correc = erlarg;
            ertest = tko_max(epsabs, epsrel * fabs(reseps));
            if (abserr <= ertest) break;
        }
        if (numrl2 == 1) noext = 1;
        if (ifail == 5) break;
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = 0;
        small = small * 5.0e-01;
        erlarg = errsum;
        goto L90;
L80:
        small = fabs(b - a) * 3.75e-01;
        erlarg = errsum;
        ertest = errbnd;
        rlist2[2] = area;
L90:
        ;
    }
    /* a Fortran DO that runs to completion leaves last = limit+1; it cannot happen here
     * because last == limit always sets ifail = 1 and exits (or jumps to 115) */

    if (abserr == TKO_HUGE) goto L115;
    if (ifail + ierro == 0) goto L110;
    if (ierro == 3) abserr = abserr + correc;
    if (ifail == 0) ifail = 3;
    if (result != 0.0 && area != 0.0) goto L105;
    if (abserr > errsum) goto L115;
    if (area == 0.0) goto L130;
    goto L110;
L105:
    if (abserr / fabs(result) > errsum / fabs(area)) goto L115;
L110:
    if (ksgn == (-1) && tko_max(fabs(result), fabs(area)) <= defabs * 1.0e-02) goto L130;
    if (1.0e-02 > (result / area) || (result / area) > 1.0e+02 || errsum > fabs(area)) ifail = 6;
    goto L130;
L115:
    result = 0.0;
    for (k = 1; k <= last; k++) result = result + rlist[k];
    abserr = errsum;
L130:
    if (ifail > 2) ifail = ifail - 1;
L140:
    *result_out = result;
    *abserr_out = abserr;
    if (neval_out) *neval_out = 42 * last - 21;
    if (ifail_out) *ifail_out = ifail;
    if (last_out) *last_out = last;
}

/* integration.f90:231-244 (iqags with its default tolerances) */
static double iqags(tko_f1 f, void *ctx, double a, double b, int *last, int *ifail)
{
    double res, abserr;
    int neval;
    tko_qags(f, ctx, a, b, 1.0e-12, 1.0e-06, &res, &abserr, &neval, ifail, last);
    return res;
}

/* integration.f90:51-64 */
double tko_adaptive_integrate(tko_f1 f, void *ctx, double x, double xi, int last[3], int ifail[3])
{
    (void)xi;
    const double eps = 1e-15;
    const double ymin = -1.0 + eps;
    const double ymax = 1.0 - eps;
    int l[3], e[3];
    double integral = 0.0;
    integral = integral + iqags(f, ctx, ymin, -fabs(x) - eps, &l[0], &e[0]);
    integral = integral + iqags(f, ctx, -fabs(x) + eps, fabs(x) - eps, &l[1], &e[1]);
    integral = integral + iqags(f, ctx, fabs(x) + eps, ymax, &l[2], &e[2]);
    if (last) { last[0] = l[0]; last[1] = l[1]; last[2] = l[2]; }
    if (ifail) { ifail[0] = e[0]; ifail[1] = e[1]; ifail[2] = e[2]; }
    return integral;
}

/* integration.f90:18-49 */
typedef struct { tko_f1 f; void *ctx; double xi; int n_pixels, grid_type; } pb_ctx;
static double pulled_back_func(double eta, void *vctx)
{
    pb_ctx *c = (pb_ctx *)vctx;
    double y = tko_push_forward(eta, c->xi, c->n_pixels, c->grid_type);
    double J = tko_push_jacob(eta, c->xi, c->n_pixels, c->grid_type);
    return c->f(y, c->ctx) * J;
}
double tko_integrate(tko_f1 f, void *ctx, double x, double xi, int n_pixels, int grid_type)
{
    double aa[8], ii[7], abserr, resabs, resasc;
    pb_ctx c = { f, ctx, xi, n_pixels, grid_type };
    aa[1] = tko_pull_back(-1.0, xi, n_pixels, grid_type);
    aa[2] = tko_pull_back(-fabs(x), xi, n_pixels, grid_type);
    aa[3] = tko_pull_back(-fabs(xi), xi, n_pixels, grid_type);
    aa[4] = tko_pull_back(0.0, xi, n_pixels, grid_type);
    aa[5] = tko_pull_back(fabs(xi), xi, n_pixels, grid_type);
    aa[6] = tko_pull_back(fabs(x), xi, n_pixels, grid_type);
    aa[7] = tko_pull_back(1.0, xi, n_pixels, grid_type);
    if (aa[3] < aa[2]) { double z = aa[2]; aa[2] = aa[3]; aa[3] = z; }
    if (aa[6] < aa[5]) { double z = aa[5]; aa[5] = aa[6]; aa[6] = z; }
    for (int i = 1; i <= 6; i++)
        tko_qk21(pulled_back_func, &c, aa[i], aa[i + 1], &ii[i], &abserr, &resabs, &resasc);
    double integral = 0.0;
    for (int i = 1; i <= 6; i++) integral = integral + ii[i];
    return integral;
}
