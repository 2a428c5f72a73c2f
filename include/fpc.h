/* fpc.h -- C-ABI of the B200-native batched four-player-chess environment.
 *
 * This is the drop-in boundary underneath the reference's pybind11 module `alphazero_cpp`
 * (/root/reference/src/cpp/wrapper.cpp:15-254).  Every entry point names the reference
 * interface it replaces.  Plain pointers and sizes only; no torch types; no exceptions cross
 * this ABI: functions return FPC_OK or a negative code and fpc_last_error() holds the text
 * (the binding turns codes into RuntimeError, wrapper.cpp:17-27).
 *
 * All `d_` pointers are DEVICE pointers on the current CUDA device; `h_` pointers are HOST
 * pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls
 * taking a stream are asynchronous with respect to the host.  There is NO CPU fallback: with no
 * usable CUDA device every compute entry point fails with FPC_ERR_CUDA.
 *
 * Geometry.  The reference fixes the board at compile time (engine/board.h:22-24).  Here the
 * side length R selects a compiled instantiation: R=14 (invalid_area 3, the 4pchess board of
 * BASELINE.json), R=8 (invalid_area 2, the reference as checked in), R=10 (2), R=13 (3).
 *
 * Board record (one game), fpc_record_bytes(R) bytes, 16-byte aligned in arrays:
 *   [0, R*R)      piece bytes, row-major, the reference's Piece bits (engine/board.h:101-104):
 *                 present<<7 | color<<5 | type<<2; empty = 0x18
 *   [R*R]         side to move (0 RED, 1 BLUE, 2 YELLOW, 3 GREEN)
 *   [R*R+1, +5)   castling rights per colour (engine/board.h:290-291): 0x80 | ks<<6 | qs<<5
 *   [R*R+5, +9)   king square per colour (row*R+col; R*R = captured).  Output only: kernels
 *                 recompute it from the squares.
 *   rest          zero padding.   R=14: 208 B, R=8: 80 B.
 *
 * Move: the reference's 8-byte chess::Move image (engine/board.h:419-435) as a little-endian
 * uint64: byte0 from, byte1 to, byte2 captured Piece bits (0x18 none), byte3 promotion type
 * (6 none), byte4/5 rook from/to (R*R none), byte6 rights before, byte7 rights after (0 absent).
 *
 * Canonical move order.  The reference's move order depends on call history;
 * every list returned here is sorted by (flat action index, promotion type).
 */
#ifndef FPC_H_
#define FPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPC_MAX_MOVES 300 /* engine/board.h:706 move_buffer_size_ */
#define FPC_NUM_STATE_CHANNELS 24 /* src/cpp/board.h:21 */

#define FPC_OK 0
#define FPC_ERR_ARG (-1)      /* bad argument (unsupported R, null pointer, negative n) */
#define FPC_ERR_CUDA (-2)     /* CUDA runtime error, or no device */
#define FPC_ERR_MOVE (-3)     /* "piece missing for move" (engine/board.cpp:1046-1054) */
#define FPC_ERR_OVERFLOW (-4) /* more than FPC_MAX_MOVES pseudo-legal moves (reference aborts) */

/* status word written per game by fpc_observe / fpc_playout_step */
#define FPC_STATUS_RESULT_MASK 0x3     /* GameResult: 0 IN_PROGRESS 1 WIN_RY 2 WIN_BG 3 STALEMATE */
#define FPC_STATUS_IN_CHECK 0x100      /* side to move is in check (set only when it has no legal move) */
#define FPC_STATUS_CAN_TAKE_KING 0x200 /* some legal move captures a king */
#define FPC_STATUS_OVERFLOW 0x400      /* move buffer overflow */
#define FPC_STATUS_FINISHED 0x800      /* playout only: the game in this slot ended and was re-seeded */
#define FPC_STATUS_CHECK 0x1000        /* side to move is in check (engine/board.cpp:941-960), legal moves or not */

const char *fpc_last_error(void);
int fpc_version(void);

/* statics of fpchess::Board (src/cpp/board.cpp:9-14, wrapper.cpp:175-180) */
int fpc_supported(int R);           /* 1 if R is a compiled geometry */
int fpc_invalid_area(int R);        /* Board::invalidArea */
int fpc_record_bytes(int R);
int fpc_num_action_channels(int R); /* Board::num_action_channels = 8R+8 */
int fpc_action_space_size(int R);   /* Board::action_space_size */
int fpc_state_space_size(int R);    /* Board::state_space_size = 24*R*R */

/* fpchess::Move index map (src/cpp/move.cpp:23-104): host-side, no device needed. */
uint64_t fpc_move_from_flat(int R, int flat_index); /* Move(int flat_index), move.cpp:41-61 */
int fpc_move_flat_index(int R, uint64_t move);      /* Move::GetFlatIndex, -1 where GetIndex throws */

/* FEN -> board record (host-side, no device needed): the 4pchess-style FEN of src/py/start_fens.py as
 * src/py/fen_parser.py:104-170 parses it -- fields split on '-': field 0 the side to move (R/B/Y/G), fields 2 / 3
 * kingside / queenside castling availability "a,b,c,d" in colour order, last field the placement (rows split on
 * '/', cells on ','; "rP" = red pawn ..., "x" = one skipped cell, an integer = that many empty cells).  The
 * reference's Python path computes the castling rights and then drops them (fen_parser.py:137-140,170):
 * honour_castling = 0 reproduces that (all rights off), 1 keeps the FEN's rights.  h_record receives
 * fpc_record_bytes(R) bytes.  Returns FPC_ERR_ARG with the parser's message in fpc_last_error() on bad input. */
int fpc_record_from_fen(int R, const char *fen, int honour_castling, uint8_t *h_record);

/* ---- device-pointer batch operations ------------------------------------------------------ */

/* Observation of a batch (rules_kernel, plus expand_kernel when a dense tensor is asked for; FPC_FLAG_INCREMENTAL is
 * ignored here -- it needs a handle, see fpc_observe_tracked).  d_planes / d_mask
 * must be 16-byte aligned.  For each of n board records:
 *   legal moves      = fpchess::Board::GetLegalMoves (src/cpp/board.cpp:94-118) over
 *                      chess::Board::GetPseudoLegalMoves2 (engine/board.cpp:846-889)
 *   status           = chess::Board::GetGameResult (engine/board.cpp:891-939), order-independent
 *                      contract, plus the FPC_STATUS_* flags
 *   planes           = Board::GetEncodedStates (src/cpp/board.cpp:305-356): [n][24][R][R] f32,
 *                      rotated by k quarter turns (d_k[g] if d_k != NULL, else k_all; k_all = -1
 *                      means "each game by its own side to move")
 *   mask             = FourPlayerChess.get_legal_moves_mask
 *                      (src/py/four_player_chess_board.py:36-55): [n][8R+8][R][R] f32, unrotated
 *   flat indices     = Board::GetLegalMovesIndices (src/cpp/board.cpp:424-449) as flat indices
 * Any output pointer may be NULL.  d_moves / d_flat are [n][FPC_MAX_MOVES]. */
int fpc_observe(int R, const uint8_t *d_boards, int n, uint64_t *d_moves, int32_t *d_flat, int32_t *d_counts,
                int32_t *d_status, float *d_planes, const int32_t *d_k, int k_all, float *d_mask, int flags,
                void *stream);

/* Dense outputs and streams.  The compact results (moves, counts, status, boards) are written by
 * rules_kernel on `stream`.  The dense f32 planes / mask are written by expand_kernel on an
 * internal second stream as soon as the rules kernel has produced the records of their ones; by default
 * `stream` then waits for it, so everything is ordered on `stream` as usual.  With
 * FPC_FLAG_ASYNC_DENSE that wait is left out: the next call's rules kernel overlaps this call's
 * expansion (the HBM-bound part), and the caller orders a consumer of the dense tensors with
 * fpc_join(consumer_stream).  State is per host thread and device. */
#define FPC_FLAG_ASYNC_DENSE 1
int fpc_join(void *stream);

/* FPC_FLAG_INCREMENTAL: resident dense tensors updated in place.  A fpc_dense_track handle owns the record of the
 * cells the latest dense call THROUGH IT set to 1.0 in one planes tensor and / or one mask tensor (created for a
 * board size and batch size on the current device).  When a *_tracked call carries this flag and the handle knows
 * the content of every tensor the call asks for, the rules kernel clears the previously set cells and sets the
 * new ones -- some 120 scattered 4-byte stores per game instead of rewriting 113 KB -- and no expansion runs.
 * The result is bit-identical to a full rewrite PROVIDED nothing else wrote to the tensors since the previous call
 * through the handle; the caller declares any such write with fpc_dense_track_invalidate().  Otherwise (first use,
 * another tensor pointer, after an invalidate, no handle) the call does the full rewrite.  Nothing is inferred from
 * pointer identity alone: a handle only vouches for tensors it has itself written since it was created or
 * invalidated, so a tensor freed and re-allocated at the same address must come with a new (or invalidated) handle.
 * Calls through one handle must be issued on one stream, or ordered by the caller.  Everything is ordered on `stream`. */
#define FPC_FLAG_INCREMENTAL 2
typedef struct fpc_dense_track fpc_dense_track;
fpc_dense_track *fpc_dense_track_create(int R, int n); /* NULL on failure (fpc_last_error) */
void fpc_dense_track_invalidate(fpc_dense_track *track);
void fpc_dense_track_destroy(fpc_dense_track *track);   /* waits for the device; NULL is allowed */
/* fpc_observe / fpc_playout_step (below) with a handle; track == NULL is the plain call. */
int fpc_observe_tracked(fpc_dense_track *track, int R, const uint8_t *d_boards, int n, uint64_t *d_moves, int32_t *d_flat,
                        int32_t *d_counts, int32_t *d_status, float *d_planes, const int32_t *d_k, int k_all, float *d_mask,
                        int flags, void *stream);
int fpc_playout_step_tracked(fpc_dense_track *track, int R, uint8_t *d_boards, int n, uint64_t seed, uint64_t *d_game,
                             int32_t *d_ply, const uint8_t *d_start, int max_plies, uint64_t game_stride, uint64_t *d_chosen,
                             int32_t *d_counts, int32_t *d_status, float *d_planes, const int32_t *d_k, int k_all,
                             float *d_mask, uint64_t *d_counters, int flags, void *stream);

/* Releases what the calling host thread holds on the current device (internal streams, events, record workspace),
 * after waiting for its pending work.  The next dense call re-creates them. */
int fpc_shutdown(void);

/* Measurement hook: time every expand_kernel and (dense-path) rules_kernel launch with CUDA events on the
 * streams they run on (up to 4096 launches per enable).  fpc_profile_read waits for those streams and
 * returns the number of timed launches and the summed durations in milliseconds.  The instrumentation is not free:
 * four timestamped event records per step drain the two internal streams between launches and lengthen a 74 us step
 * to 79 us (measured on B200, tools/overlap_probe.py FPC_P_NOPROF), so throughput is measured with the hook off. */
int fpc_profile_enable(int on);
int fpc_profile_read(int *launches, double *expand_ms, double *rules_ms);

/* Board::GetEncodedStates alone (no move generation). */
int fpc_encode(int R, const uint8_t *d_boards, int n, const int32_t *d_k, int k_all, float *d_planes,
               int flags, void *stream);

/* chess::Board::MakeMove (engine/board.cpp:1028-1096) with full generator moves, as used by
 * Board::TakeAction (src/cpp/board.cpp:234-239).  d_err[g] = FPC_OK or FPC_ERR_MOVE.
 * d_in == d_out is allowed. */
int fpc_make_moves(int R, const uint8_t *d_in, const uint64_t *d_moves, int n, uint8_t *d_out, int32_t *d_err,
                   void *stream);

/* MakeMove with index-built moves Move(flat_index) (src/cpp/move.cpp:41-61) -- the self-play
 * path (src/cpp/node.cpp:87-92, src/py/alphazero.py:119-121): only from/to, so no promotion,
 * no rook move, no rights update. */
int fpc_make_index(int R, const uint8_t *d_in, const int32_t *d_flat, int n, uint8_t *d_out, int32_t *d_err,
                   void *stream);

/* chess::Board::CalculateHeuristic (engine/board.cpp:1263-1292) for the team to move. */
int fpc_heuristic(int R, const uint8_t *d_boards, int n, int32_t *d_value, void *stream);

/* One ply of the deterministic random playout for every game slot (BASELINE.json configs[1]):
 * observe the position (planes/mask optional, as fpc_observe), then either finish the game
 * (result != IN_PROGRESS, or d_ply[g]+1 == max_plies) and re-seed the slot from h/d start
 * record with game id += game_stride, or play
 * legal[ ((mix(seed, game, ply) >> 32) * n_legal) >> 32 ] with MakeMove(full).
 * d_boards is updated in place.  d_counters (8 x uint64, may be NULL) accumulates:
 * [0] positions, [1] finished games, [2..5] result histogram of finished games (index =
 * GameResult, 0 = ply cap), [6] sum of n_legal, [7] overflow count. */
int fpc_playout_step(int R, uint8_t *d_boards, int n, uint64_t seed, uint64_t *d_game, int32_t *d_ply,
                     const uint8_t *d_start, int max_plies, uint64_t game_stride, uint64_t *d_chosen,
                     int32_t *d_counts, int32_t *d_status, float *d_planes, const int32_t *d_k, int k_all,
                     float *d_mask, uint64_t *d_counters, int flags, void *stream);

/* Viewer queries (src/cpp/board.cpp:120-232; the pygame UI only).  d_out [n][R*R] bytes, one per square in row-major
 * order (every row / column, the cut corners included, as the reference's loops run):
 *   bit c (0..3) = fpchess::Board::IsAttackedByPlayer(square, colour c) (src/cpp/board.cpp:142-209: rays bounded by the
 *                  R x R index range only, so they run through the cut corners; a piece of another colour or a
 *                  non-slider of the same colour blocks)
 *   bit 4 + t    = chess::Board::IsAttackedByTeam(team t, square) (engine/board.cpp:606-786) */
int fpc_attack_maps(int R, const uint8_t *d_boards, int n, uint8_t *d_out, void *stream);

/* ---- batched PUCT: one warp owns one game's tree in HBM ---------------------------------------
 *
 * Replaces fpchess::Node (src/cpp/node.{h,cpp}) driven by MCTS.search / step / expand
 * (src/py/mcts.py:17-89).  The tree of game g lives in slab g of caller-owned device arrays
 * (struct-of-arrays, node_cap entries per game, children of a node contiguous and in ascending flat
 * action index = torch.nonzero order, mcts.py:83-87).  Node 0 is the root.  Boards are materialised
 * only for nodes that were selected as a leaf (board_cap per game; one per simulation plus the
 * root): a child's board is its parent's board + MakeMove(Move(flat_index)) (node.cpp:87-92), made
 * when the child is first reached instead of at expansion -- same boards, fewer copies.
 *
 * One simulation for all games = fpc_tree_select -> network forward on `d_planes` (PyTorch) ->
 * fpc_tree_expand_backup.
 */
typedef struct fpc_tree {
  int32_t R, n_games, node_cap, board_cap;
  double C;               /* Node::C (alphazero.py:294) */
  /* per node, [n_games][node_cap] */
  int32_t *parent;        /* -1 for the root */
  int32_t *first_child;   /* index of the first child in the same slab */
  int32_t *n_children;    /* 0 = not expanded (Node::IsExpanded, node.cpp:14-17) */
  int32_t *visits;        /* Node::visit_count */
  int32_t *move_flat;     /* flat action index of Node::move_made */
  int32_t *board_idx;     /* slot in `boards`, -1 = not materialised */
  double *value_sum;      /* Node::value_sum */
  float *prior;           /* Node::prior (a float32 policy value widened to double by the reference) */
  /* per game, [n_games] */
  int32_t *n_nodes, *n_boards;
  int32_t *leaf;          /* node chosen by the last select, -1 = none (root dropped) */
  int32_t *dropped;       /* 1 once a terminal leaf was reached: the root sits out the remaining
                             simulations (mcts.py:22-23) */
  int32_t *error;         /* bit 0 node_cap exceeded, bit 1 board_cap exceeded, bit 2 no child selectable
                             (the reference throws, node.cpp:72-75), bit 3 index-built move failed */
  uint8_t *boards;        /* [n_games][board_cap][record] */
  /* the leaf batch: input of the rules kernel and of the network */
  uint8_t *leaf_boards;   /* [n_games][record] */
  int32_t *leaf_flat;     /* [n_games][FPC_MAX_MOVES] legal flat indices of the leaf, ascending */
  int32_t *leaf_counts;   /* [n_games] */
  int32_t *leaf_status;   /* [n_games] FPC_STATUS_* of the leaf */
  int32_t *k;             /* [n_games] quarter turns the leaf's planes were rotated by */
} fpc_tree;

/* Roots: node 0 of every game = d_root_boards[g], visit_count 1 (mcts.py:30), no children. */
int fpc_tree_reset(const fpc_tree *t, const uint8_t *d_root_boards, void *stream);

/* Node::ChooseLeaf (node.cpp:19-47) for every live game: descend with Node::SelectChild
 * (node.cpp:49-78: argmax of Q + C*sqrt(ln(sqrt(N))/(1+n))*P in double, first maximum wins),
 * materialise the leaf's board, run the rules kernel on the leaf batch (legal moves + GetGameResult)
 * and encode it into d_planes [n_games][24][R][R] (Board::GetEncodedStates).  batch_rotation != 0
 * reproduces the reference: every leaf is rotated by the colour of the first live leaf
 * (src/cpp/board.cpp:354-355; the first leaf that is not terminal, as mcts.py:18-26 drops those before encoding);
 * 0 rotates each leaf by its own side to move.  flags + track (may be NULL): FPC_FLAG_INCREMENTAL updates d_planes
 * in place (the network only reads it between two selects), see above. */
int fpc_tree_select(const fpc_tree *t, int batch_rotation, float *d_planes, int flags, fpc_dense_track *track, void *stream);

/* MCTS.step after the network (mcts.py:66-79) + MCTS.expand (mcts.py:82-89) for every game with a
 * leaf: terminal leaf -> Backpropagate(0 | -1) and drop the root (node.cpp:33-43); otherwise
 * softmax(d_logits[g]) -> ParseActionspace (rotate back by the leaf's k) -> * legal mask ->
 * renormalise -> Node::BackpropagateNodes(value) -> one child per non-zero prior (visit_count 1,
 * node.h:28).  d_logits [n_games][A*R*R] f32 in the network's (rotated) frame, d_values [n_games]. */
int fpc_tree_expand_backup(const fpc_tree *t, const float *d_logits, const float *d_values, void *stream);

/* ---- host-buffer operations (per-object calls of the binding; e2e measurement) ------------- */

typedef struct fpc_ctx fpc_ctx;

/* One context per (device, R): a stream plus pinned staging and device buffers for up to
 * max_n boards.  Not thread-safe; one per host thread. */
fpc_ctx *fpc_ctx_create(int device, int R, int max_n);
void fpc_ctx_destroy(fpc_ctx *ctx);
void *fpc_ctx_stream(fpc_ctx *ctx);

/* Host in, host out.  Any output may be NULL.  h_planes / h_mask are host buffers; when the
 * caller wants the tensors left on the device (the reference's device="cuda"), it passes
 * d_planes / d_mask (device pointers) instead and NULL for the host ones. */
int fpc_host_observe(fpc_ctx *ctx, const uint8_t *h_boards, int n, uint64_t *h_moves, int32_t *h_flat,
                     int32_t *h_counts, int32_t *h_status, float *h_planes, float *d_planes, int k_all,
                     float *h_mask, float *d_mask);
int fpc_host_make_moves(fpc_ctx *ctx, const uint8_t *h_in, const uint64_t *h_moves, int n, uint8_t *h_out,
                        int32_t *h_err);
int fpc_host_make_index(fpc_ctx *ctx, const uint8_t *h_in, const int32_t *h_flat, int n, uint8_t *h_out,
                        int32_t *h_err);
/* One playout ply through host buffers: boards, game ids and plies are copied in, stepped and
 * copied back; planes/mask stay on the device (d_planes/d_mask may be NULL).  The call returns when
 * the host buffers hold the results.  With FPC_FLAG_ASYNC_DENSE it does not wait for the dense
 * tensors: their expansion overlaps the next call, and fpc_ctx_sync() waits for it. */
int fpc_host_playout_step(fpc_ctx *ctx, uint8_t *h_boards, int n, uint64_t seed, uint64_t *h_game,
                          int32_t *h_ply, const uint8_t *h_start, int max_plies, uint64_t game_stride,
                          int32_t *h_counts, int32_t *h_status, float *d_planes, int k_all, float *d_mask,
                          int flags);
int fpc_ctx_sync(fpc_ctx *ctx);

/* Run the context's copies and kernels on `stream` (e.g. PyTorch's current stream, when outputs go into tensors that
 * stream's allocator owns) instead of the context's own one; NULL switches back.  Host-buffer calls still return only
 * when their results have landed. */
int fpc_ctx_set_stream(fpc_ctx *ctx, void *stream);
int fpc_current_device(void); /* cudaGetDevice, or a negative code */

/* Page-locked host memory for the h_ buffers above (pageable memory works too, through the driver's staging). */
void *fpc_host_alloc(size_t bytes); /* NULL on failure (fpc_last_error) */
void fpc_host_free(void *p);

/* Node::ExpandNodes (src/cpp/node.cpp:79-131) / Board::TakeAction (src/cpp/board.cpp:234-239) for a batch in ONE trip:
 * child[i] = MakeMove(h_parents[i], h_moves[i]) (full 8-byte moves), and the observation of every child -- status word
 * (GetGameResult + flags) and legal-move count -- so that the binding answers the children's GetGameResult /
 * GetLegalMoves without another trip.  The children's legal moves stay in the context ([n][FPC_MAX_MOVES] on the
 * device) until the next call through it; fpc_host_fetch_moves copies the first `stride` moves of each of the n rows
 * to h_moves [n][stride] (stride = the largest count the caller needs).  h_flat (may be NULL) receives the flat
 * action indices the same way when asked for in fpc_host_fetch_moves. */
int fpc_host_expand(fpc_ctx *ctx, const uint8_t *h_parents, const uint64_t *h_moves, int n, uint8_t *h_children,
                    int32_t *h_err, int32_t *h_counts, int32_t *h_status);
int fpc_host_fetch_moves(fpc_ctx *ctx, int n, int stride, uint64_t *h_moves);
/* Board::GetEncodedStates alone from host records into a device tensor (k_all as in fpc_encode). */
int fpc_host_encode(fpc_ctx *ctx, const uint8_t *h_boards, int n, int k_all, float *d_planes);
int fpc_host_heuristic(fpc_ctx *ctx, const uint8_t *h_boards, int n, int32_t *h_value);
int fpc_host_attack_maps(fpc_ctx *ctx, const uint8_t *h_boards, int n, uint8_t *h_out);

/* ---- environment-owned board store + DLPack -------------------------------------------------------
 *
 * For consumers that are not PyTorch: the library owns the device memory of n games (board records, per-game
 * observation outputs, the dense planes / mask tensors) and hands any of them out as a DLPack DLManagedTensor
 * (dlpack.h v0.8 layout, restated below so that this header stands alone), zero-copy.  The managed tensor keeps the
 * environment alive: the store is freed when fpc_env_destroy has been called AND every exported tensor's deleter has
 * run.  All work of an environment is ordered on its own stream (fpc_env_stream); fpc_env_sync waits for it. */
typedef struct fpc_env fpc_env;
fpc_env *fpc_env_create(int device, int R, int n); /* NULL on failure (fpc_last_error) */
void fpc_env_destroy(fpc_env *env);
void *fpc_env_stream(fpc_env *env);
int fpc_env_sync(fpc_env *env);
int fpc_env_set_boards(fpc_env *env, const uint8_t *h_boards, int first, int count); /* host records -> store */
int fpc_env_get_boards(fpc_env *env, uint8_t *h_boards, int first, int count);
/* fpc_observe over the store into the environment's own outputs (which = bit set of FPC_ENV_* below) */
int fpc_env_observe(fpc_env *env, int which, int k_all);
/* fpc_playout_step over the store (start record: host pointer), same outputs */
int fpc_env_playout_step(fpc_env *env, uint64_t seed, const uint8_t *h_start, int max_plies, uint64_t game_stride,
                         int which, int k_all);
#define FPC_ENV_BOARDS 1   /* uint8  [n][record]            */
#define FPC_ENV_COUNTS 2   /* int32  [n]                    */
#define FPC_ENV_STATUS 4   /* int32  [n]                    */
#define FPC_ENV_MOVES 8    /* int64  [n][FPC_MAX_MOVES]     */
#define FPC_ENV_FLAT 16    /* int32  [n][FPC_MAX_MOVES]     */
#define FPC_ENV_PLANES 32  /* float32 [n][24][R][R]         */
#define FPC_ENV_MASK 64    /* float32 [n][8R+8][R][R]       */
#define FPC_ENV_PLY 128    /* int32  [n]   (playout)        */
#define FPC_ENV_GAME 256   /* int64  [n]   (playout)        */

/* dlpack.h (v0.8) restated: identical layout, guarded so that a translation unit including the real header first
 * keeps its definitions. */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { DLDeviceType device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void *data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t *shape;
  int64_t *strides; /* NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void *manager_ctx;
  void (*deleter)(struct DLManagedTensor *self);
} DLManagedTensor;
#endif
/* One of the FPC_ENV_* tensors as a DLManagedTensor (the consumer calls ->deleter when done; wrap it in a PyCapsule
 * named "dltensor" for torch.utils.dlpack.from_dlpack / cupy.from_dlpack).  NULL on failure. */
DLManagedTensor *fpc_env_dlpack(fpc_env *env, int which);

#ifdef __cplusplus
}
#endif
#endif /* FPC_H_ */
