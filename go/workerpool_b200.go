//go:build cgo

// Package gubernator: drop-in replacement for WorkerPool (workers.go:54-61) that evaluates rate limits on a B200
// through libgubernator_b200.so.  NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no Go toolchain is present there);
// it is written against the reference's exact method set and kept small enough to review by eye.  The same C ABI is
// exercised by the C++ host layer (gubernator_b200/csrc/host_v1.cpp) and the Python tests.
//
// Wiring (one line in NewV1Instance, gubernator.go:129):
//
//	s.workerPool = NewB200WorkerPool(&conf)   // instead of NewWorkerPool(&conf)
//
// Build with CGO_ENABLED=1 (the reference Dockerfile sets 0, Dockerfile:22) and
// CGO_LDFLAGS="-L<repo>/gubernator_b200 -lgubernator_b200".
package gubernator

/*
#cgo CFLAGS: -I${SRCDIR}/../include
#cgo LDFLAGS: -lgubernator_b200
#include <stdlib.h>
#include "gubernator_b200.h"
*/
import "C"

import (
	"context"
	"fmt"
	"sync"
	"time"
	"unsafe"

	"github.com/OneOfOne/xxhash"
	"github.com/mailgun/holster/v4/clock"
	"github.com/pkg/errors"
	"github.com/segmentio/fasthash/fnv1"
)

// B200WorkerPool has the method set V1Instance uses on WorkerPool: GetRateLimit (gubernator.go:598, global.go:245),
// Load (:143), Store (:161), AddCacheItem (:452), GetCacheItem, Close (:169).
type B200WorkerPool struct {
	conf  *Config
	table *C.gub_table

	mu      sync.Mutex
	pending []*pendingReq // requests waiting for the next device batch
	timer   *time.Timer
	// pinned request/response arenas (gub_host_alloc), two sets so one fills while the other is in flight
	reqs  [2]unsafe.Pointer
	resps [2]unsafe.Pointer
	cur   int
	// key strings by fingerprint, kept only when a Loader/Store needs CacheItem.Key back (store.go:49-78)
	keys map[[2]uint64]string
}

type pendingReq struct {
	req   *RateLimitReq
	state RateLimitReqState
	done  chan pendingResp
}
type pendingResp struct {
	resp *RateLimitResp
	err  error
}

const (
	b200MaxBatch  = 65536                  // device batch; RPCs (<= 1000 items each, gubernator.go:40) are coalesced up to this
	b200BatchWait = 500 * time.Microsecond // same window as BehaviorConfig.BatchWait (config.go:128)
)

func NewB200WorkerPool(conf *Config) *B200WorkerPool {
	cfg := C.gub_config{
		capacity_slots: C.uint64_t(2 * conf.CacheSize), // load factor <= 0.5; replaces CacheSize/Workers LRUs (workers.go:132)
		max_batch:      b200MaxBatch,
		device:         0,
	}
	p := &B200WorkerPool{conf: conf}
	if rc := C.gub_create(&cfg, &p.table); rc != 0 {
		panic(fmt.Sprintf("gub_create: %s", C.GoString(C.gub_last_error())))
	}
	for i := range p.reqs {
		p.reqs[i] = C.gub_host_alloc(C.size_t(b200MaxBatch * 64))
		p.resps[i] = C.gub_host_alloc(C.size_t(b200MaxBatch * 32))
	}
	if conf.Loader != nil || conf.Store != nil {
		p.keys = make(map[[2]uint64]string)
	}
	return p
}

// GetRateLimit mirrors WorkerPool.GetRateLimit (workers.go:261-290): the caller blocks until its request has been
// evaluated.  Instead of a channel hop to a worker goroutine, the request joins the next device batch.
func (p *B200WorkerPool) GetRateLimit(ctx context.Context, r *RateLimitReq, st RateLimitReqState) (*RateLimitResp, error) {
	pr := &pendingReq{req: r, state: st, done: make(chan pendingResp, 1)}
	p.mu.Lock()
	p.pending = append(p.pending, pr)
	if len(p.pending) >= b200MaxBatch {
		batch := p.takeLocked()
		p.mu.Unlock()
		p.flush(batch)
	} else {
		if p.timer == nil {
			p.timer = time.AfterFunc(b200BatchWait, p.flushTimer)
		}
		p.mu.Unlock()
	}
	select {
	case out := <-pr.done:
		return out.resp, out.err
	case <-ctx.Done():
		return nil, ctx.Err() // workers.go:278-288
	}
}

func (p *B200WorkerPool) takeLocked() []*pendingReq {
	b := p.pending
	p.pending = nil
	if p.timer != nil {
		p.timer.Stop()
		p.timer = nil
	}
	return b
}

func (p *B200WorkerPool) flushTimer() {
	p.mu.Lock()
	b := p.takeLocked()
	p.mu.Unlock()
	if len(b) > 0 {
		p.flush(b)
	}
}

// flush evaluates one batch: fill gub_req records, one cgo call, hand every caller its response.
func (p *B200WorkerPool) flush(batch []*pendingReq) {
	p.mu.Lock() // one batch in flight per arena; the device serialises batches anyway
	arena := p.cur
	p.cur ^= 1
	p.mu.Unlock()
	reqs := unsafe.Slice((*C.gub_req)(p.reqs[arena]), b200MaxBatch)
	resps := unsafe.Slice((*C.gub_resp)(p.resps[arena]), b200MaxBatch)
	for i, pr := range batch {
		r := pr.req
		key := r.HashKey() // client.go:39-41
		q := &reqs[i]
		q.key_xxh64 = C.uint64_t(xxhash.ChecksumString64S(key, 0)) // workers.go:153-155
		q.key_fnv1 = C.uint64_t(fnv1.HashString64(key))            // replicated_hash.go:108
		q.hits, q.limit, q.duration, q.burst = C.int64_t(r.Hits), C.int64_t(r.Limit), C.int64_t(r.Duration), C.int64_t(r.Burst)
		q.created_at = C.int64_t(*r.CreatedAt) // defaulted by GetRateLimits (gubernator.go:218-220)
		q.algorithm = C.uint32_t(r.Algorithm)
		q.behavior = C.uint32_t(r.Behavior) & 0xff
		if pr.state.IsOwner {
			q.behavior |= C.GUB_REQ_IS_OWNER
		}
		if p.keys != nil {
			p.keys[[2]uint64{uint64(q.key_xxh64), uint64(q.key_fnv1) >> 8}] = key
		}
		if r.Algorithm == Algorithm_LEAKY_BUCKET && r.Burst == 0 {
			r.Burst = r.Limit // the reference mutates the request (algorithms.go:264-266); keep that visible to callers
		}
	}
	var clk C.gub_clock
	C.gub_clock_fill(C.int64_t(clock.Now().UnixNano()/1000000), &clk)
	rc := C.gub_submit(p.table, &reqs[0], C.size_t(len(batch)), &clk, &resps[0])
	for i, pr := range batch {
		if rc != 0 {
			pr.done <- pendingResp{nil, errors.New(C.GoString(C.gub_last_error()))}
			continue
		}
		o := &resps[i]
		if o.err_code != 0 {
			pr.done <- pendingResp{nil, b200Error(int(o.err_code), pr.req)}
			continue
		}
		if o.status == C.GUB_OVER_LIMIT && pr.state.IsOwner {
			metricOverLimitCounter.Add(1) // algorithms.go:164,184,242; also available in bulk from gub_get_counters
		}
		pr.done <- pendingResp{&RateLimitResp{
			Status: Status(o.status), Limit: int64(o.limit), Remaining: int64(o.remaining), ResetTime: int64(o.reset_time),
		}, nil}
	}
}

// b200Error rebuilds the error values handleGetRateLimit returns (workers.go:299-320); V1Instance wraps them further.
func b200Error(code int, r *RateLimitReq) error {
	scope := "Error in tokenBucket"
	if r.Algorithm == Algorithm_LEAKY_BUCKET {
		scope = "Error in leakyBucket"
	}
	switch code {
	case C.GUB_ERR_INVALID_ALGORITHM:
		return errors.Errorf("Invalid rate limit algorithm '%d'", r.Algorithm)
	case C.GUB_ERR_GREGORIAN_WEEKS:
		return errors.Wrap(errors.New("`Duration = GregorianWeeks` not yet supported; consider making a PR!`"), scope)
	case C.GUB_ERR_GREGORIAN_INVALID:
		return errors.Wrap(errors.New("behavior DURATION_IS_GREGORIAN is set; but `Duration` is not a valid gregorian interval"), scope)
	case C.GUB_ERR_TABLE_FULL:
		return errors.New("rate limit table is full")
	}
	return errors.Errorf("device error %d", code)
}

func toItem(key string, item *CacheItem) (C.gub_item, bool) {
	var it C.gub_item
	it.key_xxh64 = C.uint64_t(xxhash.ChecksumString64S(key, 0))
	it.key_fnv1 = C.uint64_t(fnv1.HashString64(key))
	it.expire_at = C.int64_t(item.ExpireAt)
	switch v := item.Value.(type) {
	case *TokenBucketItem:
		it.algorithm = C.GUB_TOKEN_BUCKET
		it.status, it.limit, it.duration, it.remaining, it.stamp = C.int32_t(v.Status), C.int64_t(v.Limit), C.int64_t(v.Duration), C.int64_t(v.Remaining), C.int64_t(v.CreatedAt)
	case *LeakyBucketItem:
		it.algorithm = C.GUB_LEAKY_BUCKET
		it.limit, it.duration, it.remaining_f, it.stamp, it.burst = C.int64_t(v.Limit), C.int64_t(v.Duration), C.double(v.Remaining), C.int64_t(v.UpdatedAt), C.int64_t(v.Burst)
	default:
		return it, false
	}
	return it, true
}

func (p *B200WorkerPool) fromItem(it *C.gub_item) *CacheItem {
	item := &CacheItem{ExpireAt: int64(it.expire_at), Algorithm: Algorithm(it.algorithm)}
	if p.keys != nil {
		item.Key = p.keys[[2]uint64{uint64(it.key_xxh64), uint64(it.key_fnv1) >> 8}]
	}
	if it.algorithm == C.GUB_LEAKY_BUCKET {
		item.Value = &LeakyBucketItem{Limit: int64(it.limit), Duration: int64(it.duration), Remaining: float64(it.remaining_f), UpdatedAt: int64(it.stamp), Burst: int64(it.burst)}
	} else {
		item.Value = &TokenBucketItem{Status: Status(it.status), Limit: int64(it.limit), Duration: int64(it.duration), Remaining: int64(it.remaining), CreatedAt: int64(it.stamp)}
	}
	return item
}

// AddCacheItem mirrors workers.go:537 (used by UpdatePeerGlobals, gubernator.go:452).
func (p *B200WorkerPool) AddCacheItem(ctx context.Context, key string, item *CacheItem) error {
	it, ok := toItem(key, item)
	if !ok {
		return nil
	}
	if p.keys != nil {
		p.mu.Lock()
		p.keys[[2]uint64{uint64(it.key_xxh64), uint64(it.key_fnv1) >> 8}] = key
		p.mu.Unlock()
	}
	if C.gub_add_items(p.table, &it, 1) != 0 {
		return errors.New(C.GoString(C.gub_last_error()))
	}
	return nil
}

// GetCacheItem mirrors workers.go:583.
func (p *B200WorkerPool) GetCacheItem(ctx context.Context, key string) (*CacheItem, bool, error) {
	kx := C.uint64_t(xxhash.ChecksumString64S(key, 0))
	kf := C.uint64_t(fnv1.HashString64(key))
	var it C.gub_item
	var found C.uint8_t
	if C.gub_get_items(p.table, &kx, &kf, 1, C.int64_t(MillisecondNow()), &it, &found) != 0 {
		return nil, false, errors.New(C.GoString(C.gub_last_error()))
	}
	if found == 0 {
		return nil, false, nil
	}
	item := p.fromItem(&it)
	item.Key = key
	return item, true, nil
}

// Load mirrors workers.go:329: stream Loader.Load() into the table in bulk.
func (p *B200WorkerPool) Load(ctx context.Context) error {
	ch, err := p.conf.Loader.Load()
	if err != nil {
		return errors.Wrap(err, "Error in loader.Load")
	}
	buf := make([]C.gub_item, 0, 65536)
	flush := func() error {
		if len(buf) == 0 {
			return nil
		}
		if C.gub_add_items(p.table, &buf[0], C.size_t(len(buf))) != 0 {
			return errors.New(C.GoString(C.gub_last_error()))
		}
		buf = buf[:0]
		return nil
	}
	for item := range ch {
		if it, ok := toItem(item.Key, item); ok {
			p.keys[[2]uint64{uint64(it.key_xxh64), uint64(it.key_fnv1) >> 8}] = item.Key
			buf = append(buf, it)
			if len(buf) == cap(buf) {
				if err := flush(); err != nil {
					return err
				}
			}
		}
	}
	return flush()
}

// Store mirrors workers.go:451: scan the table and hand every item to Loader.Save().
func (p *B200WorkerPool) Store(ctx context.Context) error {
	var n C.size_t
	if C.gub_size(p.table, &n) != 0 {
		return errors.New(C.GoString(C.gub_last_error()))
	}
	items := make([]C.gub_item, int(n)+1)
	if C.gub_scan(p.table, &items[0], C.size_t(len(items)), &n) != 0 {
		return errors.New(C.GoString(C.gub_last_error()))
	}
	out := make(chan *CacheItem, 500)
	go func() {
		for i := 0; i < int(n) && i < len(items); i++ {
			out <- p.fromItem(&items[i])
		}
		close(out)
	}()
	return p.conf.Loader.Save(out)
}

// Close mirrors workers.go:157.
func (p *B200WorkerPool) Close() error {
	for i := range p.reqs {
		C.gub_host_free(p.reqs[i])
		C.gub_host_free(p.resps[i])
	}
	C.gub_destroy(p.table)
	return nil
}
