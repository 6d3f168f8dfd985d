/* group_tmpl.h -- curve group + MSM template of the CPU oracle (TEST INFRASTRUCTURE).
 *   #define K(name)  coordinate field prefix (Fp or Fp2 instance)
 *   #define FR(name) scalar field prefix     #define FR_BITS 254|255|...     #define FR_NL its 64-bit limbs (4; 6 for bw6-761)
 *   #define IMPL_C_OK(c) the window widths the reference implements for this curve (implementedCs, multiexp.go:77)
 *   #define GP(name) group prefix
 * Restates (reference tree, ecc/bn254; ecc/bls12-381 is the same generated code):
 *   g1.go:822-985 addMixed/subMixed/doubleMixed/doubleNegMixed, :736-817 add/double,
 *   :726-731 unsafeFromJacExtended, :150-166 FromJacobian,
 *   multiexp.go:32-146 MultiExp (bestC :75-93, split :103-140), :148-209 _innerMsmG1,
 *   :302-315 msmReduceChunk, :681-693 computeNbChunks/lastC, :709-803 partitionScalars,
 *   multiexp_jacobian.go:8-61 processChunkG1Jacobian.
 */
typedef struct { K(t) x, y; } GP(aff);
typedef struct { K(t) x, y, zz, zzz; } GP(xyzz);
typedef struct { K(t) x, y, z; } GP(jac);

static inline int GP(aff_is_inf)(const GP(aff)* a) { return K(is_zero)(&a->x) && K(is_zero)(&a->y); }
static inline void GP(xyzz_set_inf)(GP(xyzz)* p) { K(set_one)(&p->x); K(set_one)(&p->y); K(set_zero)(&p->zz); K(set_zero)(&p->zzz); }

/* doubleMixed (negate=0) g1.go:962-985 / doubleNegMixed (negate=1) :933-957 */
static void GP(double_mixed)(GP(xyzz)* p, const GP(aff)* a, int negate) {
  K(t) U, V, W, S, XX, M, S2, L, y;
  if (negate) K(neg)(&y, &a->y); else y = a->y;
  K(dbl)(&U, &y);
  K(sqr)(&V, &U);
  K(mul)(&W, &U, &V);
  K(mul)(&S, &a->x, &V);
  K(sqr)(&XX, &a->x);
  K(dbl)(&M, &XX);
  K(add)(&M, &M, &XX);
  K(dbl)(&S2, &S);
  K(mul)(&L, &W, &y);
  K(sqr)(&p->x, &M);
  K(sub)(&p->x, &p->x, &S2);
  K(sub)(&p->y, &S, &p->x);
  K(mul)(&p->y, &p->y, &M);
  K(sub)(&p->y, &p->y, &L);
  p->zz = V;
  p->zzz = W;
}

/* addMixed g1.go:822-873 / subMixed :878-930 */
static void GP(add_mixed)(GP(xyzz)* p, const GP(aff)* a, int negate) {
  if (GP(aff_is_inf)(a)) return;
  K(t) ay;
  if (negate) K(neg)(&ay, &a->y); else ay = a->y;
  if (K(is_zero)(&p->zz)) {
    p->x = a->x;
    p->y = ay;
    K(set_one)(&p->zz);
    K(set_one)(&p->zzz);
    return;
  }
  K(t) P, R;
  K(mul)(&P, &a->x, &p->zz);
  K(sub)(&P, &P, &p->x);
  K(mul)(&R, &ay, &p->zzz);
  K(sub)(&R, &R, &p->y);
  if (K(is_zero)(&P)) {
    if (K(is_zero)(&R)) { GP(double_mixed)(p, a, negate); return; }
    K(set_zero)(&p->zz);
    K(set_zero)(&p->zzz);
    return;
  }
  K(t) PP, PPP, Q, Q2, RR, X3, Y3;
  K(sqr)(&PP, &P);
  K(mul)(&PPP, &P, &PP);
  K(mul)(&Q, &p->x, &PP);
  K(sqr)(&RR, &R);
  K(sub)(&X3, &RR, &PPP);
  K(dbl)(&Q2, &Q);
  K(sub)(&p->x, &X3, &Q2);
  K(sub)(&Y3, &Q, &p->x);
  K(mul)(&Y3, &Y3, &R);
  K(mul)(&R, &p->y, &PPP);
  K(sub)(&p->y, &Y3, &R);
  K(mul)(&p->zz, &p->zz, &PP);
  K(mul)(&p->zzz, &p->zzz, &PPP);
}

/* double g1.go:795-817 */
static void GP(xyzz_double)(GP(xyzz)* p, const GP(xyzz)* q) {
  K(t) U, V, W, S, XX, M;
  GP(xyzz) r;
  K(dbl)(&U, &q->y);
  K(sqr)(&V, &U);
  K(mul)(&W, &U, &V);
  K(mul)(&S, &q->x, &V);
  K(sqr)(&XX, &q->x);
  K(dbl)(&M, &XX);
  K(add)(&M, &M, &XX);
  K(mul)(&U, &W, &q->y);
  K(sqr)(&r.x, &M);
  K(sub)(&r.x, &r.x, &S);
  K(sub)(&r.x, &r.x, &S);
  K(sub)(&r.y, &S, &r.x);
  K(mul)(&r.y, &r.y, &M);
  K(sub)(&r.y, &r.y, &U);
  K(mul)(&r.zz, &V, &q->zz);
  K(mul)(&r.zzz, &W, &q->zzz);
  *p = r;
}

/* add g1.go:736-788 */
static void GP(xyzz_add)(GP(xyzz)* p, const GP(xyzz)* q) {
  if (K(is_zero)(&q->zz)) return;
  if (K(is_zero)(&p->zz)) { *p = *q; return; }
  K(t) A, B, U1, U2, S1, S2;
  K(mul)(&U2, &q->x, &p->zz);
  K(mul)(&U1, &p->x, &q->zz);
  K(sub)(&A, &U2, &U1);
  K(mul)(&S2, &q->y, &p->zzz);
  K(mul)(&S1, &p->y, &q->zzz);
  K(sub)(&B, &S2, &S1);
  if (K(is_zero)(&A)) {
    if (K(is_zero)(&B)) { GP(xyzz_double)(p, q); return; }
    K(set_zero)(&p->zz);
    K(set_zero)(&p->zzz);
    return;
  }
  K(t) PP, PPP, Q, V;
  K(sqr)(&PP, &A);
  K(mul)(&PPP, &A, &PP);
  K(mul)(&Q, &U1, &PP);
  K(mul)(&V, &S1, &PPP);
  K(sqr)(&p->x, &B);
  K(sub)(&p->x, &p->x, &PPP);
  K(sub)(&p->x, &p->x, &Q);
  K(sub)(&p->x, &p->x, &Q);
  K(sub)(&p->y, &Q, &p->x);
  K(mul)(&p->y, &p->y, &B);
  K(sub)(&p->y, &p->y, &V);
  K(mul)(&p->zz, &p->zz, &q->zz);
  K(mul)(&p->zz, &p->zz, &PP);
  K(mul)(&p->zzz, &p->zzz, &q->zzz);
  K(mul)(&p->zzz, &p->zzz, &PPP);
}

/* unsafeFromJacExtended g1.go:726-731 (infinity -> (0,0,0)) then FromJacobian :150-166 */
static void GP(xyzz_to_jac)(GP(jac)* j, const GP(xyzz)* p) {
  if (K(is_zero)(&p->zz)) { K(set_zero)(&j->x); K(set_zero)(&j->y); K(set_zero)(&j->z); return; }
  K(sqr)(&j->x, &p->zz);
  K(mul)(&j->x, &j->x, &p->x);
  K(sqr)(&j->y, &p->zzz);
  K(mul)(&j->y, &j->y, &p->y);
  j->z = p->zzz;
}
static void GP(jac_to_aff)(GP(aff)* a, const GP(jac)* j) {
  if (K(is_zero)(&j->z)) { K(set_zero)(&a->x); K(set_zero)(&a->y); return; }
  K(t) ai, b;
  K(inv)(&ai, &j->z);
  K(sqr)(&b, &ai);
  K(mul)(&a->x, &j->x, &b);
  K(mul)(&a->y, &j->y, &b);
  K(mul)(&a->y, &a->y, &ai);
}
static void GP(xyzz_to_aff)(GP(aff)* a, const GP(xyzz)* p) {
  GP(jac) j;
  GP(xyzz_to_jac)(&j, p);
  GP(jac_to_aff)(a, &j);
}

/* ---- windows ---- */
static inline int GP(nb_chunks)(int c) { return (FR_BITS + c - 1) / c; }
static inline int GP(last_c)(int c) { return c + 1 - (GP(nb_chunks)(c) * c - FR_BITS); }
static int GP(best_c)(size_t n) { /* multiexp.go:75-93 */
  double mn = 1e300;
  int C = 4;
  for (int c = 4; c <= 16; c++) {
    if (!(IMPL_C_OK(c))) continue;
    double cc = (double)(FR_BITS + 1) * (double)(n + ((size_t)1 << c));
    double cost = cc / (double)c;
    if (cost < mn) { mn = cost; C = c; }
  }
  return C;
}

/* partitionScalars multiexp.go:709-803 for scalars [lo, hi); digits[chunk*n + i], n = total stride */
static void GP(partition_range)(const FR(t)* scalars, size_t lo, size_t hi, size_t n, int c, uint32_t* digits) {
  const int W = GP(nb_chunks)(c);
  const uint64_t mask = ((uint64_t)1 << c) - 1;
  const int64_t mx = ((int64_t)1 << (c - 1)) - 1;
  for (size_t i = lo; i < hi; i++) {
    for (int ch = 0; ch < W; ch++) digits[(size_t)ch * n + i] = 0;
    if (FR(is_zero)(&scalars[i])) continue;
    FR(t) k;
    FR(from_mont)(&k, &scalars[i]);
    int64_t carry = 0;
    for (int ch = 0; ch < W; ch++) {
      uint64_t jc = (uint64_t)ch * c, idx = jc / 64, shift = jc - idx * 64;
      int64_t d = carry + (int64_t)((k.l[idx] & (mask << shift)) >> shift);
      int multi = (64 % c != 0) && shift > (uint64_t)(64 - c) && idx < FR_NL - 1;
      if (multi) {
        uint64_t nb_hi = shift - (64 - c);
        d += (int64_t)((k.l[idx + 1] & (((uint64_t)1 << nb_hi) - 1)) << (c - nb_hi));
      }
      if (ch < W - 1) {
        carry = 0;
        if (d > mx) { d -= (int64_t)1 << c; carry = 1; }
        if (d == 0) continue;
        digits[(size_t)ch * n + i] = d > 0 ? ((uint32_t)d << 1) : ((((uint32_t)(-d - 1)) << 1) + 1);
      } else {
        digits[(size_t)ch * n + i] = (uint32_t)d << 1;
      }
    }
  }
}

/* processChunkG1Jacobian multiexp_jacobian.go:8-61 */
static void GP(process_chunk)(int c_eff, const GP(aff)* points, const uint32_t* digits, size_t n, GP(xyzz)* total) {
  size_t nb = (size_t)1 << (c_eff - 1);
  GP(xyzz)* buckets = (GP(xyzz)*)malloc(nb * sizeof(GP(xyzz)));
  for (size_t k = 0; k < nb; k++) GP(xyzz_set_inf)(&buckets[k]);
  for (size_t i = 0; i < n; i++) {
    uint32_t e = digits[i];
    if (e == 0) continue;
    if ((e & 1) == 0) GP(add_mixed)(&buckets[(e >> 1) - 1], &points[i], 0);
    else GP(add_mixed)(&buckets[e >> 1], &points[i], 1);
  }
  GP(xyzz) run, tot;
  GP(xyzz_set_inf)(&run);
  GP(xyzz_set_inf)(&tot);
  for (size_t k = nb; k-- > 0;) {
    if (!K(is_zero)(&buckets[k].zz)) GP(xyzz_add)(&run, &buckets[k]);
    GP(xyzz_add)(&tot, &run);
  }
  *total = tot;
  free(buckets);
}

/* batchAddG1Affine g1.go:1122-1182: R[j] += P[j] for cnt independent affine pairs (no infinity, distinct x)
 * with one shared inversion (Montgomery's trick) */
static void GP(batch_add_affine)(GP(aff)** R, const GP(aff)* P, int cnt, K(t)* lambda, K(t)* lambdain) {
  if (cnt == 0) return;
  for (int j = 0; j < cnt; j++) K(sub)(&lambdain[j], &P[j].x, &R[j]->x);
  K(t) acc;
  K(set_one)(&lambda[0]);
  acc = lambdain[0];
  for (int i = 1; i < cnt; i++) { lambda[i] = acc; K(mul)(&acc, &acc, &lambdain[i]); }
  K(inv)(&acc, &acc);
  for (int i = cnt - 1; i > 0; i--) { K(mul)(&lambda[i], &lambda[i], &acc); K(mul)(&acc, &acc, &lambdain[i]); }
  lambda[0] = acc;
  for (int j = 0; j < cnt; j++) {
    K(t) t;
    GP(aff) Q;
    K(sub)(&t, &P[j].y, &R[j]->y);
    K(mul)(&lambda[j], &lambda[j], &t);
    K(sqr)(&Q.x, &lambda[j]);
    K(sub)(&Q.x, &Q.x, &R[j]->x);
    K(sub)(&Q.x, &Q.x, &P[j].x);
    K(sub)(&t, &R[j]->x, &Q.x);
    K(mul)(&Q.y, &lambda[j], &t);
    K(sub)(&Q.y, &Q.y, &R[j]->y);
    *R[j] = Q;
  }
}

/* batch sizes of the generated processors, multiexp_affine.go:298-338 (0 = no batch-affine processor) */
static inline int GP(batch_size)(int c) {
  switch (c) { case 10: return 80; case 11: return 150; case 12: return 200; case 13: return 350; case 14: return 400;
               case 15: return 500; case 16: return 640; default: return 0; }
}

/* processChunkG1BatchAffine multiexp_affine.go:24-231: affine buckets, batches of independent additions sharing
 * one inversion, a queue for conflicting buckets, extended-Jacobian fallback buckets for doublings / a full queue */
typedef struct { uint32_t bucket; GP(aff) point; } GP(batch_op);
static void GP(process_chunk_batch_affine)(int c_eff, int batch, const GP(aff)* points, const uint32_t* digits, size_t n,
                                           GP(xyzz)* total) {
  const size_t nb = (size_t)1 << (c_eff - 1);
  GP(aff)* buckets = (GP(aff)*)calloc(nb, sizeof(GP(aff)));          /* infinity = (0,0) */
  GP(xyzz)* bucketsJE = (GP(xyzz)*)malloc(nb * sizeof(GP(xyzz)));
  for (size_t k = 0; k < nb; k++) GP(xyzz_set_inf)(&bucketsJE[k]);
  unsigned char* in_batch = (unsigned char*)calloc(nb, 1);            /* bitSet */
  GP(aff)** R = (GP(aff)**)malloc(sizeof(GP(aff)*) * batch);
  GP(aff)* P = (GP(aff)*)malloc(sizeof(GP(aff)) * batch);
  uint32_t* batch_ids = (uint32_t*)malloc(sizeof(uint32_t) * batch);
  GP(batch_op)* queue = (GP(batch_op)*)malloc(sizeof(GP(batch_op)) * batch);
  K(t)* lam = (K(t)*)malloc(sizeof(K(t)) * 2 * batch);
  int cpt = 0, qid = 0;
#define EXECUTE_AND_RESET() do { GP(batch_add_affine)(R, P, cpt, lam, lam + batch); \
    for (int u_ = 0; u_ < cpt; u_++) { in_batch[batch_ids[u_]] = 0; } \
    cpt = 0; } while (0)
  for (size_t i = 0; i < n; i++) {
    const uint32_t e = digits[i];
    if (e == 0 || GP(aff_is_inf)(&points[i])) continue;
    uint32_t b = e >> 1;
    const int is_add = (e & 1) == 0;
    if (is_add) b -= 1;
    GP(aff) pt = points[i];
    if (!is_add) K(neg)(&pt.y, &pt.y);
    if (in_batch[b]) {                       /* conflict: queue it (multiexp_affine.go:179-193) */
      queue[qid].bucket = b; queue[qid].point = pt; qid++;
      if (qid == batch - 1) { for (int q = 0; q < qid; q++) GP(add_mixed)(&bucketsJE[queue[q].bucket], &queue[q].point, 0); qid = 0; }
      continue;
    }
    /* add(): special cases without the batch (:108-146) */
    GP(aff)* BK = &buckets[b];
    if (GP(aff_is_inf)(BK)) { *BK = pt; continue; }
    if (K(eq)(&BK->x, &pt.x)) {
      if (K(eq)(&BK->y, &pt.y)) GP(add_mixed)(&bucketsJE[b], &pt, 0);   /* doubling: rare, other bucket set */
      else { K(set_zero)(&BK->x); K(set_zero)(&BK->y); }                  /* P + (-P) */
      continue;
    }
    in_batch[b] = 1; batch_ids[cpt] = b; R[cpt] = BK; P[cpt] = pt; cpt++;
    if (cpt == batch) {
      EXECUTE_AND_RESET();
      for (int q = qid - 1; q >= 0; q--) {   /* processTopQueue :155-164 */
        if (in_batch[queue[q].bucket]) break;
        GP(aff)* QB = &buckets[queue[q].bucket];
        GP(aff) qp = queue[q].point;
        qid--;
        if (GP(aff_is_inf)(QB)) { *QB = qp; continue; }
        if (K(eq)(&QB->x, &qp.x)) {
          if (K(eq)(&QB->y, &qp.y)) GP(add_mixed)(&bucketsJE[queue[q].bucket], &qp, 0);
          else { K(set_zero)(&QB->x); K(set_zero)(&QB->y); }
          continue;
        }
        in_batch[queue[q].bucket] = 1; batch_ids[cpt] = queue[q].bucket; R[cpt] = QB; P[cpt] = qp; cpt++;
      }
    }
  }
  EXECUTE_AND_RESET();
  for (int q = 0; q < qid; q++) GP(add_mixed)(&bucketsJE[queue[q].bucket], &queue[q].point, 0);
#undef EXECUTE_AND_RESET
  GP(xyzz) run, tot;
  GP(xyzz_set_inf)(&run);
  GP(xyzz_set_inf)(&tot);
  for (size_t k = nb; k-- > 0;) {            /* :212-221 */
    GP(add_mixed)(&run, &buckets[k], 0);
    if (!K(is_zero)(&bucketsJE[k].zz)) GP(xyzz_add)(&run, &bucketsJE[k]);
    GP(xyzz_add)(&tot, &run);
  }
  *total = tot;
  free(buckets); free(bucketsJE); free(in_batch); free(R); free(P); free(batch_ids); free(queue); free(lam);
}

/* msmReduceChunk multiexp.go:302-315 */
static void GP(reduce_chunks)(int c, int W, const GP(xyzz)* totals, GP(xyzz)* out) {
  GP(xyzz) acc = totals[W - 1];
  for (int j = W - 2; j >= 0; j--) {
    for (int l = 0; l < c; l++) GP(xyzz_double)(&acc, &acc);
    GP(xyzz_add)(&acc, &totals[j]);
  }
  *out = acc;
}

/* ---- threaded MSM: the reference's decomposition (recursive halving + one task per window) on a
 * worker pool ---- */
typedef struct {
  size_t off, n;
  int c, W;
  uint32_t* digits;
  GP(xyzz)* totals;
} GP(leaf);

typedef struct {
  const GP(aff)* points;
  const FR(t)* scalars;
  GP(leaf)* leaves;
  int nleaves;
  /* phase 1: digit tasks (leaf, range); phase 2: window tasks (leaf, window) */
  int phase;
  int use_batch_affine;
  size_t ntasks;
  size_t next;
  pthread_mutex_t mu;
  size_t* task_leaf;
  size_t* task_a;
  size_t* task_b;
} GP(job);

static void* GP(worker)(void* arg) {
  GP(job)* jb = (GP(job)*)arg;
  for (;;) {
    pthread_mutex_lock(&jb->mu);
    size_t t = jb->next++;
    pthread_mutex_unlock(&jb->mu);
    if (t >= jb->ntasks) break;
    GP(leaf)* lf = &jb->leaves[jb->task_leaf[t]];
    if (jb->phase == 1) {
      GP(partition_range)(jb->scalars + lf->off, jb->task_a[t], jb->task_b[t], lf->n, lf->c, lf->digits);
    } else {
      int j = (int)jb->task_a[t];
      int ce = (j == lf->W - 1) ? GP(last_c)(lf->c) : lf->c;
      const uint32_t* dg = lf->digits + (size_t)j * lf->n;
      /* getChunkProcessorG1 (multiexp.go:213-299): batch-affine for 10 <= c <= 16 when at least batchSize buckets of the
       * window are hit (chunkStat.nbBucketFilled, multiexp.go:807-853), extended Jacobian otherwise */
      int batch = jb->use_batch_affine ? GP(batch_size)(ce) : 0;
      if (batch) {
        size_t nbk = (size_t)1 << (ce - 1), filled = 0;
        unsigned char* seen = (unsigned char*)calloc(nbk, 1);
        for (size_t i = 0; i < lf->n && filled < (size_t)batch; i++) {
          uint32_t e = dg[i];
          if (!e) continue;
          uint32_t b = (e >> 1) - ((e & 1) ? 0 : 1);
          if (!seen[b]) { seen[b] = 1; filled++; }
        }
        free(seen);
        if (filled < (size_t)batch) batch = 0;
      }
      if (batch) GP(process_chunk_batch_affine)(ce, batch, jb->points + lf->off, dg, lf->n, &lf->totals[j]);
      else GP(process_chunk)(ce, jb->points + lf->off, dg, lf->n, &lf->totals[j]);
    }
  }
  return NULL;
}

static void GP(run_pool)(GP(job)* jb, int nthreads) {
  jb->next = 0;
  if (nthreads <= 1) { GP(worker)(jb); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, GP(worker), jb);
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  free(th);
}

/* MultiExp's split recursion, multiexp.go:95-140 */
static int GP(cost_function)(int nb_tasks, int nb_cpus, long cost_per_task) {
  long total = nb_tasks;
  while (nb_tasks >= nb_cpus) { nb_tasks -= nb_cpus; total += cost_per_task; }
  if (nb_tasks > 0) total += cost_per_task;
  return (int)(total > 0x7fffffff ? 0x7fffffff : total);
}
static void GP(collect_leaves)(size_t off, size_t n, int nb_tasks, int force_c, GP(leaf)* leaves, int* nleaves, int max_leaves) {
  int C = force_c ? force_c : GP(best_c)(n);
  if (!force_c && n >= 2 && *nleaves + 2 <= max_leaves) {
    int nbc = GP(nb_chunks)(C);
    long pre = GP(cost_function)(nbc, nb_tasks, (long)n + (1L << C));
    int c2 = GP(best_c)(n / 2);
    long post = GP(cost_function)(GP(nb_chunks)(c2) * 2, nb_tasks, (long)(n / 2) + (1L << c2));
    if (post < pre) {
      int half_tasks = (nb_tasks + 1) / 2;
      GP(collect_leaves)(off, n / 2, half_tasks, 0, leaves, nleaves, max_leaves);
      GP(collect_leaves)(off + n / 2, n - n / 2, half_tasks, 0, leaves, nleaves, max_leaves);
      return;
    }
  }
  GP(leaf)* lf = &leaves[(*nleaves)++];
  lf->off = off; lf->n = n; lf->c = C; lf->W = GP(nb_chunks)(C);
  lf->digits = NULL; lf->totals = NULL;
}

/* result: affine normal form in out_aff, raw Jacobian (xyzz -> jac of the combined sum) in out_jac */
static int GP(msm)(const GP(aff)* points, const FR(t)* scalars, size_t n, int force_c, int nthreads, int nb_tasks,
                   GP(aff)* out_aff, GP(jac)* out_jac, int* used_c, int* used_leaves, int use_batch_affine) {
  if (nthreads < 1) nthreads = 1;
  if (nb_tasks <= 0) nb_tasks = 2 * nthreads; /* multiexp.go:67-68 with NumCPU := nthreads */
  GP(xyzz) sum;
  GP(xyzz_set_inf)(&sum);
  if (n > 0) {
    enum { MAXL = 64 };
    GP(leaf) leaves[MAXL];
    int nl = 0;
    GP(collect_leaves)(0, n, nb_tasks, force_c, leaves, &nl, MAXL);
    size_t maxtasks = 0;
    for (int i = 0; i < nl; i++) {
      leaves[i].digits = (uint32_t*)malloc(sizeof(uint32_t) * leaves[i].n * (size_t)leaves[i].W + 4);
      leaves[i].totals = (GP(xyzz)*)malloc(sizeof(GP(xyzz)) * leaves[i].W);
      maxtasks += (size_t)leaves[i].W + (size_t)nthreads * 4;
    }
    GP(job) jb;
    jb.points = points; jb.scalars = scalars; jb.leaves = leaves; jb.nleaves = nl;
    jb.use_batch_affine = use_batch_affine;
    pthread_mutex_init(&jb.mu, NULL);
    jb.task_leaf = (size_t*)malloc(sizeof(size_t) * maxtasks);
    jb.task_a = (size_t*)malloc(sizeof(size_t) * maxtasks);
    jb.task_b = (size_t*)malloc(sizeof(size_t) * maxtasks);
    /* phase 1: digits, parallel over scalars (parallel.Execute, multiexp.go:741) */
    size_t nt = 0;
    for (int i = 0; i < nl; i++) {
      size_t parts = (size_t)nthreads * 4;
      if (parts > leaves[i].n) parts = leaves[i].n ? leaves[i].n : 1;
      for (size_t p = 0; p < parts; p++) {
        jb.task_leaf[nt] = i;
        jb.task_a[nt] = leaves[i].n * p / parts;
        jb.task_b[nt] = leaves[i].n * (p + 1) / parts;
        nt++;
      }
    }
    jb.phase = 1; jb.ntasks = nt;
    GP(run_pool)(&jb, nthreads);
    /* phase 2: one task per (leaf, window), high window first (multiexp.go:180-206) */
    nt = 0;
    for (int i = 0; i < nl; i++)
      for (int j = leaves[i].W - 1; j >= 0; j--) { jb.task_leaf[nt] = i; jb.task_a[nt] = j; jb.task_b[nt] = 0; nt++; }
    jb.phase = 2; jb.ntasks = nt;
    GP(run_pool)(&jb, nthreads);
    for (int i = 0; i < nl; i++) {
      GP(xyzz) r;
      GP(reduce_chunks)(leaves[i].c, leaves[i].W, leaves[i].totals, &r);
      GP(xyzz_add)(&sum, &r);  /* join of the halves (AddAssign, multiexp.go:136-139) */
      free(leaves[i].digits);
      free(leaves[i].totals);
    }
    if (used_c) *used_c = leaves[0].c;
    if (used_leaves) *used_leaves = nl;
    free(jb.task_leaf); free(jb.task_a); free(jb.task_b);
    pthread_mutex_destroy(&jb.mu);
  }
  GP(jac) j;
  GP(xyzz_to_jac)(&j, &sum);
  if (out_jac) *out_jac = j;
  if (out_aff) GP(jac_to_aff)(out_aff, &j);
  return 0;
}

/* [k]P, k an FR_NL-limb canonical integer, double-and-add */
static void GP(scalar_mul)(GP(aff)* out, const GP(aff)* p, const uint64_t* k) {
  GP(xyzz) acc;
  GP(xyzz_set_inf)(&acc);
  for (int i = 64 * FR_NL - 1; i >= 0; i--) {
    GP(xyzz_double)(&acc, &acc);
    if ((k[i >> 6] >> (i & 63)) & 1) GP(add_mixed)(&acc, p, 0);
  }
  GP(xyzz_to_aff)(out, &acc);
}

/* out[i] = [start + i] * base for i in [0, n): chains of affine additions with one shared inversion
 * per step across `lanes` independent chains (Montgomery trick, as batchAddG1Affine g1.go:1122-1182) */
typedef struct { const GP(aff)* base; uint64_t start; size_t lo, hi; GP(aff)* out; } GP(gen_arg);
static void GP(aff_add_distinct_batch)(GP(aff)* pts, size_t cnt, const GP(aff)* d, K(t)* scratch /* 2*cnt */) {
  /* pts[i] += d, assuming pts[i] != +-d and neither is infinity (true for consecutive multiples
     away from the group order); denominators x_d - x_i inverted together */
  K(t)* den = scratch; K(t)* pref = scratch + cnt;
  K(t) acc; K(set_one)(&acc);
  for (size_t i = 0; i < cnt; i++) { K(sub)(&den[i], &d->x, &pts[i].x); pref[i] = acc; K(mul)(&acc, &acc, &den[i]); }
  K(t) inv; K(inv)(&inv, &acc);
  for (size_t i = cnt; i-- > 0;) {
    K(t) di; K(mul)(&di, &inv, &pref[i]); K(mul)(&inv, &inv, &den[i]);
    K(t) lam, x3, y3, t;
    K(sub)(&lam, &d->y, &pts[i].y); K(mul)(&lam, &lam, &di);
    K(sqr)(&x3, &lam); K(sub)(&x3, &x3, &pts[i].x); K(sub)(&x3, &x3, &d->x);
    K(sub)(&t, &pts[i].x, &x3); K(mul)(&y3, &lam, &t); K(sub)(&y3, &y3, &pts[i].y);
    pts[i].x = x3; pts[i].y = y3;
  }
}
static void* GP(gen_worker)(void* a_) {
  GP(gen_arg)* a = (GP(gen_arg)*)a_;
  size_t n = a->hi - a->lo;
  if (n == 0) return NULL;
  /* lanes chains: chain l produces indices lo + l*len + s, s = 0..len-1 */
  size_t lanes = n < 256 ? 1 : 256;
  size_t len = (n + lanes - 1) / lanes;
  GP(aff)* cur = (GP(aff)*)malloc(sizeof(GP(aff)) * lanes);
  K(t)* scratch = (K(t)*)malloc(sizeof(K(t)) * 2 * lanes);
  size_t active = 0;
  for (size_t l = 0; l < lanes; l++) {
    size_t idx = a->lo + l * len;
    if (idx >= a->hi) break;
    uint64_t k[FR_NL] = {a->start + idx};
    GP(scalar_mul)(&cur[l], a->base, k);
    active++;
  }
  for (size_t s = 0; s < len; s++) {
    size_t cnt = 0;
    for (size_t l = 0; l < active; l++) {
      size_t idx = a->lo + l * len + s;
      if (idx < a->hi) { a->out[idx] = cur[l]; cnt = l + 1; }
    }
    if (s + 1 < len) {
      /* generic path (handles infinity / doubling) when a special case could occur */
      int special = 0;
      for (size_t l = 0; l < cnt; l++)
        if (GP(aff_is_inf)(&cur[l]) || K(eq)(&cur[l].x, &a->base->x)) special = 1;
      if (!special) GP(aff_add_distinct_batch)(cur, cnt, a->base, scratch);
      else for (size_t l = 0; l < cnt; l++) {
        GP(xyzz) t; GP(xyzz_set_inf)(&t);
        GP(add_mixed)(&t, &cur[l], 0); GP(add_mixed)(&t, a->base, 0);
        GP(xyzz_to_aff)(&cur[l], &t);
      }
    }
  }
  free(cur); free(scratch);
  return NULL;
}
static void GP(generate_multiples)(const GP(aff)* base, uint64_t start, size_t n, GP(aff)* out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n / 1024 + 1) nthreads = (int)(n / 1024 + 1);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  GP(gen_arg)* args = (GP(gen_arg)*)malloc(sizeof(GP(gen_arg)) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    args[t].base = base; args[t].start = start; args[t].out = out;
    args[t].lo = n * t / nthreads; args[t].hi = n * (t + 1) / nthreads;
    if (nthreads == 1) GP(gen_worker)(&args[t]); else pthread_create(&th[t], NULL, GP(gen_worker), &args[t]);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(args);
}
