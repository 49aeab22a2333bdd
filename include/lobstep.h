/*
 * lobstep.h -- C ABI of the B200-native batched limit-order-book step.
 *
 * This is the drop-in boundary for ONE hot path of JaxMARL-HFT: the batched
 * limit-order-book environment step.  The reference has no FFI of its own (it
 * is 100 % Python/JAX); the de-facto interface is the method set the trainer
 * calls (all citations relative to /root/reference/gymnax_exchange):
 *
 *   lob_step_launch    <->  MARLEnv.step / step_env       jaxen/marl_env.py:776 / :212
 *   lob_reset_launch   <->  MARLEnv.reset / reset_env     jaxen/marl_env.py:764 / :130
 *   lob_replay_launch  <->  BaseLOBEnv.step_env           jaxen/base_env.py:189
 *                           job.scan_through_entire_array jaxob/JaxOrderBookArrays.py:736
 *                           (also the reset-state precompute, base_env.py:245-296)
 *   lob_l2_launch      <->  job.get_L2_state              jaxob/JaxOrderBookArrays.py:1232
 *   lob_rollout_launch <->  lax.scan(vmap(env.step)) over T steps with the actions given up front
 *                           jaxrl/MARL/ippo_rnn_JAXMARL.py:616-661, jaxen/Speed_test.py:165-214
 *   lob_host_replay_*  <->  the same replay through HOST buffers (bench e2e leg)
 *   lob_loader_*_launch <-> _pre_process_msg_ob + merge_market_orders
 *                           jaxlobster/lobster_loader.py:891-945, :1073-1132
 *   lob_draw_launch*   <->  the jax.random products of one step (marl_env.py:135, :294-295; base_env.py:222-225;
 *                           exec_env.py:221), counter-based, for callers without JAX
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / jax types.
 *   - every *_launch takes DEVICE pointers and a cudaStream_t (as void*); it
 *     validates on the host, enqueues on that stream, never synchronises and
 *     never allocates.  Return 0 on success, a negative LOB_E_* otherwise;
 *     lob_last_error() gives the text (thread local).
 *   - re-entrant: no mutable globals; per-device constants arrive as buffers.
 *   - int32 rows use the reference layouts verbatim:
 *       order row  i32[6] = [price, qty, order_id, trader_id, time_s, time_ns]   jaxob_constants.py:38-44
 *       trade row  i32[8] = [price, +-qty, passive_oid, aggr_oid, time_s, time_ns,
 *                            passive_tid, aggr_tid]                               jaxob_constants.py:46-54
 *       message    i32[8] = [type, side, qty, price, order_id, trader_id,
 *                            time_s, time_ns]                                     jaxob_constants.py:84-92
 *     empty rows are all -1.
 *   - state leaves are updated IN PLACE (the XLA custom call would use
 *     input_output_aliases); field order follows StatesandParams.py:14-122.
 */
#ifndef LOBSTEP_H_
#define LOBSTEP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOB_ABI_VERSION 11
#define LOB_MAX_AGENT_TYPES 8
#define LOB_MAX_AGENT_I32 4  /* int32 state leaves per agent type */
#define LOB_MAX_AGENT_F32 10 /* float32 state leaves per agent type */

/* error codes */
#define LOB_OK 0
#define LOB_E_INVALID (-1)     /* bad config / shape */
#define LOB_E_UNSUPPORTED (-2) /* valid in the reference, not built here (see DESIGN.md "next") */
#define LOB_E_CUDA (-3)        /* CUDA runtime error at launch */

/* ---- enums (values are part of the ABI) -------------------------------- */
enum LobAgentKind { LOB_AGENT_MM = 0, LOB_AGENT_EXE = 1 };

/* MarketMaking_EnvironmentConfig.action_space  jaxob_config.py:55 ; mm_env.py:161-178 */
enum LobMMActionSpace { LOB_MM_ACT_FIXED_QUANTS = 0, LOB_MM_ACT_DIRECTIONAL = 1,
                         LOB_MM_ACT_BOB_RL = 2 /* mm_env.py:1474 */, LOB_MM_ACT_BOB_STRATEGY = 3 /* mm_env.py:1400 */,
                         LOB_MM_ACT_SIMPLE = 4 /* :1123 */, LOB_MM_ACT_SPREAD_SKEW = 5 /* :1667 */, LOB_MM_ACT_AVST = 6 /* :1248 */ };
/* Execution_EnvironmentConfig.action_space     jaxob_config.py:157 ; exec_env.py:162-175 */
enum LobEXEActionSpace { LOB_EXE_ACT_FIXED_QUANTS = 0, LOB_EXE_ACT_FIXED_QUANTS_COMPLEX = 1,
                          LOB_EXE_ACT_FIXED_QUANTS_1MSG = 2 /* exec_env.py:732 */, LOB_EXE_ACT_SIMPLEST_CASE = 3 /* :935 */,
                          LOB_EXE_ACT_TWAP = 4 /* :1126 */,
                          LOB_EXE_ACT_FIXED_PRICES = 5 /* :1001; the action is a VECTOR of n_actions quantities */ };
/* observation_space  mm_env.py:2767-2788 ; exec_env.py:179-186 */
enum LobObsSpace { LOB_OBS_ENGINEERED = 0, LOB_OBS_BASIC = 1, LOB_OBS_SIMPLEST_CASE = 2 /* EXE only, exec_env.py:1841 */ };
/* MM reward_function  mm_env.py:2489-2513 */
enum LobMMReward {
  LOB_MM_REW_PORTFOLIO_VALUE = 0, LOB_MM_REW_BUY_SELL_PNL = 1, LOB_MM_REW_COMPLEX = 2,
  LOB_MM_REW_ZERO_INV = 3, LOB_MM_REW_SPOONER = 4, LOB_MM_REW_SPOONER_DAMPED = 5,
  LOB_MM_REW_SPOONER_ASYM_DAMPED = 6, LOB_MM_REW_SPOONER_ASYM_DAMPED2 = 7,
  LOB_MM_REW_SPOONER_SCALED = 8, LOB_MM_REW_DELTA_PORTFOLIO_VALUE = 9
};
/* EXE reward_function  exec_env.py:1735-1752 */
enum LobEXEReward { LOB_EXE_REW_NORMAL = 0, LOB_EXE_REW_FINISH_FAST = 1, LOB_EXE_REW_SIMPLEST_CASE = 2 };
/* reference_price / unwind_price  mm_env.py:2294-2303,2373-2396 ; exec_env.py:1564-1580 */
enum LobRefPrice { LOB_REF_MID = 0, LOB_REF_MID_AVG = 1, LOB_REF_FAR_TOUCH = 2, LOB_REF_NEAR_TOUCH = 3 };
/* inv_penalty  mm_env.py:2516-2536 */
enum LobInvPenalty { LOB_INVPEN_NONE = 0, LOB_INVPEN_LINEAR = 1, LOB_INVPEN_QUADRATIC = 2,
                     LOB_INVPEN_EXP4 = 3, LOB_INVPEN_THRESHOLD = 4 };
/* Execution task  exec_env.py:220-223 */
enum LobExeTask { LOB_TASK_RANDOM = 0, LOB_TASK_BUY = 1, LOB_TASK_SELL = 2 };

/* ---- order book configuration: JAXLOB_Configuration  jaxob_config.py:12-30 */
typedef struct LobBookConfig {
  int32_t n_orders;              /* nOrders  : rows per book side */
  int32_t n_trades;              /* nTrades  : rows of the trade log */
  int32_t maxint;                /* cfg.maxint (2147483647) */
  int32_t init_id;               /* cfg.init_id (-2) */
  int32_t book_depth;            /* cfg.book_depth */
  int32_t cancel_mode;           /* cst CancelMode: 0/1 by id then initial-liquidity match (identical in the reference:
                                    job:94-139); 2/3 add the random same-price fallbacks of job:142-164, whose uniform
                                    draws arrive in the cancel_u buffers (a PRNG product, like perm) */
  int32_t type_4_interpretation; /* 0 IOC, 1 LIM, 2 MKT  jaxob_constants.py:70-74 */
  int32_t check_book_fill;       /* job:395-401 / :484-490 */
} LobBookConfig;

/* ---- one agent type: MarketMaking_/Execution_EnvironmentConfig  jaxob_config.py:33-200 */
typedef struct LobAgentTypeConfig {
  int32_t kind;                  /* LobAgentKind */
  int32_t n_agents;              /* number_of_agents_per_type[i] */
  int32_t trader_id_start;       /* agent a has trader id trader_id_start - a   mm_env.py:193-195 */
  int32_t action_space;
  int32_t observation_space;
  int32_t reward_function;
  int32_t n_actions;
  int32_t num_messages_by_agent;        /* cancels + actions */
  int32_t num_action_messages_by_agent;
  int32_t normalize;
  int32_t time_delay_obs_act;
  int32_t fixed_quant_value;
  /* market making */
  int32_t n_ticks_offset;
  int32_t tenth_action_market_order;    /* tenth_action == "MarketOrder" */
  int32_t sell_buy_all_option;          /* mm_env.py:1018-1024 (fixed_quants), :1144-1172 (simple) */
  int32_t fixed_action_setting;
  int32_t fixed_action;
  int32_t auto_liquidate_threshold;
  int32_t unwind_price_penalty;
  int32_t inv_penalty;                  /* LobInvPenalty */
  int32_t volume_traded_bonus_market_share; /* volume_traded_bonus == "market_share"  mm_env.py:2542 */
  int32_t reference_price;              /* LobRefPrice */
  int32_t unwind_price;                 /* LobRefPrice (mid / mid_avg / far_touch) */
  int32_t clip_reward;
  int32_t exclude_extreme_spreads;
  /* execution */
  int32_t task;                         /* LobExeTask */
  int32_t task_size;
  int32_t n_ticks_in_book;
  int32_t larger_far_touch_quant;
  int32_t doom_price_penalty;
  int32_t bob_v0;                       /* market making: bobRL / bobStrategy base quantity (1, 2, 5 or 10) */
  int32_t simple_nothing_action;        /* MM 'simple': 4-entry tables (mm_env.py:1133) */
  int32_t multiplier_type_spread;       /* MM 'spread_skew': multiplier_type == "spread" (else "tick")  mm_env.py:1725 */
  /* python floats of the config: kept as double, narrowed to f32 where JAX's weak typing does */
  double auto_liquidate_alpha;
  double inv_penalty_lambda;
  double inv_penalty_quadratic_factor;
  double inv_penalty_threshold;
  double reward_scaling_quo;
  double inventoryPnL_eta;
  double inventoryPnL_gamma;
  double rebate_bps;
  double unrealizedPnL_lambda;
  double reward_lambda;
  double spread_multiplier;             /* MM 'spread_skew' */
  double skew_multiplier;
  double avst_k_parameter;              /* MM 'AvSt' */
  double avst_var_parameter;
} LobAgentTypeConfig;

/* ---- the whole step: MultiAgentConfig  jaxob_config.py:205-250 */
typedef struct LobStepConfig {
  LobBookConfig book;
  int32_t n_data_msg_per_step;   /* Nd */
  int32_t tick_size;
  int32_t ep_type_fixed_time;    /* 0 = "fixed_steps", 1 = "fixed_time" (base_env.py:358-368 message masking, the 10- /
                                    15-dim engineered observations mm_env.py:3032, exec_env.py:1943) */
  int32_t episode_time;
  int32_t order_id_counter_start;/* order_id_counter_start_when_resetting (-200) */
  int32_t placeholder_order_id;  /* -198 */
  int32_t artificial_trader_id_end_episode; /* -199 */
  int32_t artificial_order_id_end_episode;  /* -199 */
  int32_t shuffle_action_messages;
  int32_t n_agent_types;
  int32_t n_windows;             /* W: rows of the precomputed reset states */
  int32_t _pad0;
  int64_t n_messages;            /* M: rows of message_data */
  LobAgentTypeConfig agent[LOB_MAX_AGENT_TYPES];
} LobStepConfig;

/* Derived sizes (marl_env.py:85-94): n_cancel = sum n_i*(num_messages-num_action_messages),
 * n_action = sum n_i*num_action_messages, N = n_cancel + n_action + Nd. */
int32_t lob_num_msgs_per_step(const LobStepConfig* cfg);
int32_t lob_num_action_msgs(const LobStepConfig* cfg);
int32_t lob_num_cancel_msgs(const LobStepConfig* cfg);
int64_t lob_split_workspace_words(const LobStepConfig* cfg, int64_t batch);   /* size of LobStepBuffers.work_split */
int32_t lob_obs_dim(const LobStepConfig* cfg, int32_t agent_type);       /* mm_env.py:3195-3223 ; exec_env.py:2188-2202 */
int32_t lob_info_i32_cols(const LobStepConfig* cfg, int32_t agent_type);
int32_t lob_info_f32_cols(const LobStepConfig* cfg, int32_t agent_type);

/* world info columns (marl_env.py:624-639) */
#define LOB_WINFO_I32_COLS 11 /* window_index, step_counter, time_s, time_ns, order_id_counter, best_asks, best_bids,
                                 current_step, ep_done_time, abort_episode, spread */
#define LOB_WINFO_F32_COLS 4  /* end_mid_price, average_best_ask, average_best_bid, delta_time */
/* MM info columns (mm_env.py:2695-2730) */
#define LOB_MMINFO_I32_COLS 9  /* done, inventory, forced_unwind, posted_bid_price, posted_ask_price,
                                  bid_distance_from_best, ask_distance_from_best, ask_quant, bid_quant */
#define LOB_MMINFO_F32_COLS 15 /* reward, reward_portfolio_value, reward_spooner, end_of_ep_pv, reward_spooner_damped,
                                  reward_spooner_asym_damped, reward_spooner_asym_damped2, reward_delta_pv, total_PnL,
                                  delta_mid_price, market_share, buyPnL, invPnL, sellPnL, inventoryValue */
/* EXE info columns (exec_env.py:1809-1829) */
#define LOB_EXEINFO_I32_COLS 4 /* quant_left, done, doom_quant, is_sell_task */
#define LOB_EXEINFO_F32_COLS 5 /* revenue_direction_normalised, vwap_rm, drift, advantage, reward */

/* ---- buffers of one step call.  B = batch (vmap dim), N = msgs per step,
 *      No/Nt = n_orders/n_trades, n_i = agents of type i, d_i = obs dim.   */
typedef struct LobStepBuffers {
  /* WorldState leaves, in/out in place (StatesandParams.py:14-36) */
  int32_t* asks;             /* [B,No,6] ask_raw_orders */
  int32_t* bids;             /* [B,No,6] bid_raw_orders */
  int32_t* trades;           /* [B,Nt,8] */
  int32_t* init_time;        /* [B,2] */
  int32_t* window_index;     /* [B] */
  int32_t* max_steps;        /* [B] max_steps_in_episode */
  int32_t* start_index;      /* [B] */
  int32_t* step_counter;     /* [B] */
  int32_t* best_bids;        /* [B,N,2] */
  int32_t* best_asks;        /* [B,N,2] */
  int32_t* time;             /* [B,2] */
  int32_t* order_id_counter; /* [B] */
  float*   mid_price;        /* [B] */
  float*   delta_time;       /* [B] */
  /* agent state leaves per type, each [B,n_i], in/out in place (StatesandParams.py:48-74)
   *   MM : i32 {posted_distance_bid, posted_distance_ask, inventory}
   *        f32 {total_PnL, cash_balance}
   *   EXE: i32 {task_to_execute, quant_executed, is_sell_task}
   *        f32 {init_price, p_vwap, total_revenue, drift_return, advantage_return,
   *             slippage_rm, price_adv_rm, price_drift_rm, vwap_rm, trade_duration} */
  int32_t* agent_i32[LOB_MAX_AGENT_TYPES][LOB_MAX_AGENT_I32];
  float*   agent_f32[LOB_MAX_AGENT_TYPES][LOB_MAX_AGENT_F32];
  /* inputs */
  const int32_t* actions[LOB_MAX_AGENT_TYPES]; /* [B,n_i]; [B,n_i,n_actions] for the EXE fixed_prices action space */
  const int32_t* perm;             /* [B,n_action]  jax.random.permutation(shuffle_key, n_action)  marl_env.py:294-295 ; may be NULL when !shuffle */
  const int32_t* reset_window;     /* [B] randint(world_key, 0, W) of the auto-reset  base_env.py:222-225 */
  const int32_t* reset_is_sell;    /* [B,n_agent_types] randint(agent_key,0,2) per type (shared by its agents, marl_env.py:187) */
  const float*   cancel_u;         /* [B,N,2] cancel_mode 2/3 only (else may be NULL): per message the uniform [0,1) draw of
                                      jax.random.choice in get_random_id_match (job:146) and get_random_large_id_match
                                      (job:161); choice = searchsorted(cumsum(p), sum(p) * (1 - u)) */
  /* params (shared by the batch; expand_dims => no leading B) */
  const int32_t* message_data;     /* [M,8] */
  const int32_t* init_asks;        /* [W,No,6] init_states_array leaves (base_env.py:285-296) */
  const int32_t* init_bids;        /* [W,No,6] */
  const int32_t* init_trades;      /* [W,Nt,8] */
  const int32_t* init_init_time;   /* [W,2] */
  const int32_t* init_max_steps;   /* [W] */
  const int32_t* init_start_index; /* [W] */
  /* outputs */
  float*   obs[LOB_MAX_AGENT_TYPES];         /* [B,n_i,d_i]  (reset obs where done_all) */
  float*   reward[LOB_MAX_AGENT_TYPES];      /* [B,n_i] */
  uint8_t* done_all;                         /* [B] dones["__all__"] */
  uint8_t* done_agents[LOB_MAX_AGENT_TYPES]; /* [B,n_i] */
  int32_t* info_world_i32;                   /* [B,LOB_WINFO_I32_COLS] */
  float*   info_world_f32;                   /* [B,LOB_WINFO_F32_COLS] */
  int32_t* info_agent_i32[LOB_MAX_AGENT_TYPES]; /* [B,n_i,Ki] */
  float*   info_agent_f32[LOB_MAX_AGENT_TYPES]; /* [B,n_i,Kf] */
  /* optional workspace of lob_step_launch (both NULL = not used; contents are scratch, not state).  With it, books deeper
   * than 128 rows per side are stepped on a 128-row shared-memory window (the reference rests an order in the LOWEST blank
   * row, JaxOrderBookArrays.py:73, so the live orders sit in the first rows); the environments whose book does not fit
   * are listed here and redone at full capacity by a second launch of the same call.  Results are identical either way. */
  int32_t* work_redo_list;                   /* [B] environment indices of the second pass */
  int32_t* work_redo_count;                  /* [4] word 0: environments in the second pass of the LAST call; word 1: running
                                                total over all calls (statistics); 16-byte aligned */
  /* optional workspace of lob_step_launch (NULL = not used): lob_split_workspace_words(cfg, B) 32-bit words, 16-byte
   * aligned; contents are scratch, not state.  With it the step runs PIPED, as four launches of the same call -- the
   * agents' messages (one warp per environment), the message scan (books resident in shared memory), the agents' rewards /
   * state / info / observations (one thread per agent over the whole batch) and the auto-reset of the environments whose
   * episode ended -- instead of one fused kernel; books deeper than 128 rows per side keep the fused kernel and only move
   * the agents' scalar arithmetic to the per-agent launch.  Same results either way (LOB_NO_PIPE=1 forces the fused kernel). */
  int32_t* work_split;
} LobStepBuffers;

/* ---- rollout: n_steps consecutive steps of every environment in ONE call (lob_rollout_launch) -- with the work_split
 *      workspace as n_steps piped steps (4 launches each, the faster schedule), without it as ONE launch of the fused kernel,
 *      the books resident in shared memory in between -- the trainer's jit(lax.scan(vmap(env.step))) with the actions given up front
 *      (ippo_rnn_JAXMARL.py:616-661 with a pre-sampled policy, Speed_test.py:165-214).  Step ts reads row ts of the
 *      trajectory inputs and writes row ts of the trajectory outputs; a NULL input falls back to the LobStepBuffers field
 *      (the same values every step), a NULL output is not recorded.  T = n_steps, B = batch.                          */
typedef struct LobRolloutBuffers {
  int32_t n_steps;                                    /* T >= 1 */
  int32_t _pad0;
  int64_t batch;                                      /* B: the leading dimension after T of every buffer below */
  const int32_t* actions[LOB_MAX_AGENT_TYPES];        /* [T,B,n_i] ([T,B,n_i,n_actions] for EXE fixed_prices) */
  const int32_t* perm;                                /* [T,B,n_action] */
  const int32_t* reset_window;                        /* [T,B] */
  const int32_t* reset_is_sell;                       /* [T,B,n_agent_types] */
  const float*   cancel_u;                            /* [T,B,N,2] (cancel_mode 2/3) */
  float*   obs[LOB_MAX_AGENT_TYPES];                  /* [T,B,n_i,d_i] */
  float*   reward[LOB_MAX_AGENT_TYPES];               /* [T,B,n_i] */
  uint8_t* done_agents[LOB_MAX_AGENT_TYPES];          /* [T,B,n_i] */
  uint8_t* done_all;                                  /* [T,B] */
} LobRolloutBuffers;

/* ---- pure replay: every book b scans msgs[start[b] .. start[b]+n_msgs) -- */
typedef struct LobReplayBuffers {
  int32_t* asks;               /* [B,No,6] in/out */
  int32_t* bids;               /* [B,No,6] in/out */
  int32_t* trades;             /* [B,Nt,8] in/out (NOT re-initialised: base_env.py:208) */
  const int32_t* msgs;         /* [M,8] */
  const int64_t* start;        /* [B] first message of each book */
  int64_t n_msgs_total;        /* M */
  int32_t n_msgs;              /* T messages per book */
  int32_t _pad0;
  int32_t* best_out;           /* optional [B,4] = [best_ask, ask_vol, best_bid, bid_vol] after the scan (job:968), or NULL */
  const float* cancel_u;       /* [B,T,2] uniform draws for cancel_mode 2/3 (see LobStepBuffers.cancel_u), else may be NULL */
} LobReplayBuffers;

int  lob_abi_version(void);
const char* lob_last_error(void);

int lob_step_launch(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, void* cuda_stream);
int lob_reset_launch(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, void* cuda_stream);
/* roll->n_steps steps of every environment, equal to that many lob_step_launch calls fed row by row (state leaves, the
 * LobStepBuffers outputs and info hold what the LAST step left).  roll->batch must equal batch. */
int lob_rollout_launch(const LobStepConfig* cfg, const LobStepBuffers* bufs, const LobRolloutBuffers* roll, int64_t batch,
                       void* cuda_stream);
int lob_replay_launch(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books, void* cuda_stream);
/* Measurement variant of lob_replay_launch (4 books per warp; n_orders <= 112): same results, slower -- see DESIGN.md 6 */
int lob_replay_launch_grouped(const LobBookConfig* cfg, const LobReplayBuffers* bufs, int64_t n_books, void* cuda_stream);
/* asks/bids [B,No,6] -> l2 [B,4*n_levels] = [ask_p,ask_q,bid_p,bid_q] x n_levels   job:1232-1264 */
int lob_l2_launch(const LobBookConfig* cfg, const int32_t* asks, const int32_t* bids, int32_t* l2,
                  int32_t n_levels, int64_t n_books, void* cuda_stream);

/* ---- the loader's day preprocessing on the device (lobster_loader.py:891-945 _pre_process_msg_ob, :1073-1132
 * merge_market_orders).  raw = the parsed LOBSTER message table, float64 [n,6] row-major (time, type, order_id, size, price,
 * direction), time-sorted.  Call lob_loader_flags_launch, take the EXCLUSIVE prefix sum `pos` of `keep` (M = pos[n-1] +
 * keep[n-1] rows survive), then lob_loader_scatter_launch: msgs int32 [M-1,8] in the env's column order, time_out f64
 * [M-1], rows i64 [M] = original row of every kept message (book[j] = orderbook[rows[j]], j < M-1, is the book BEFORE
 * msgs[j]).  flags[0]: bit 0 = the table is not time-sorted, bit 1 = a field does not fit int32 -- both are errors. */
int lob_loader_flags_launch(const double* raw, int64_t n, int32_t day_start, int32_t day_end, int64_t* keep, int64_t* mqty,
                            int64_t* mprice, int32_t* flags, void* cuda_stream);
int lob_loader_scatter_launch(const double* raw, int64_t n, int32_t day_start, int32_t day_end, const int64_t* keep,
                              const int64_t* pos, const int64_t* mqty, const int64_t* mprice, int32_t* msgs, double* time_out,
                              int64_t* rows, int32_t* flags, void* cuda_stream);

/* The PRNG products of one step, drawn on the device by a counter-based generator: perm [B,n_action] (a uniform
 * random permutation per environment, marl_env.py:293-295), reset_window [B] in [0, n_windows) or window_selector when it
 * is >= 0 (base_env.py:222-225), reset_is_sell [B,n_agent_types] in {0,1} (exec_env.py:221) and, under cancel_mode 2/3,
 * cancel_u [B,N,2] (multiples of 2^-23 in [0,1), as jax.random.uniform yields for float32).  The reference draws these
 * with jax.random inside the step; here they are inputs of lob_step_launch, and this is the host layer's default source
 * (any other source, e.g. jax.random through the FFI stub, is equally valid).  Deterministic in (seed, counter). */
int lob_draw_launch(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, int32_t window_selector,
                    uint64_t seed, uint64_t counter, void* cuda_stream);
/* The same with the counter kept in DEVICE memory: the draw uses *counter_dev and a second one-thread kernel increments it,
 * so the call can be captured in a CUDA graph and replayed (every replay draws fresh values). */
int lob_draw_launch_dev(const LobStepConfig* cfg, const LobStepBuffers* bufs, int64_t batch, int32_t window_selector,
                        uint64_t seed, uint64_t* counter_dev, void* cuda_stream);

/* Host-buffer replay (the end-to-end leg): copies books/msgs/start host->device, replays,
 * copies books/trades back, synchronises the stream.  Device scratch is owned by the handle. */
typedef struct LobHostReplay LobHostReplay;
LobHostReplay* lob_host_replay_create(const LobBookConfig* cfg, int64_t max_books, int64_t max_msgs_total, int device);
int  lob_host_replay_set_messages(LobHostReplay* h, const int32_t* msgs_host, int64_t n_msgs_total); /* the day tensor: resident once */
int  lob_host_replay_run(LobHostReplay* h, int32_t* asks_host, int32_t* bids_host, int32_t* trades_host,
                         const int64_t* start_host, int32_t n_msgs, int64_t n_books,
                         int64_t* h2d_bytes, int64_t* d2h_bytes);
void lob_host_replay_destroy(LobHostReplay* h);

/* number of kernels this library launched on the calling thread since the last reset (bench "gpu_launches") */
int64_t lob_launch_count(void);
void    lob_launch_count_reset(void);

/* sizeof() of the interface structs, so a binding can verify its mirror of this header */
int64_t lob_sizeof_book_config(void);
int64_t lob_sizeof_agent_type_config(void);
int64_t lob_sizeof_step_config(void);
int64_t lob_sizeof_step_buffers(void);
int64_t lob_sizeof_replay_buffers(void);
int64_t lob_sizeof_rollout_buffers(void);
/* Byte offsets of sentinel fields, in the fixed order below, so that a mirror of these structs (ctypes, cgo, ...) can check
 * its LAYOUT and not only its size: LobBookConfig.{cancel_mode, check_book_fill}; LobAgentTypeConfig.{fixed_quant_value,
 * task_size, doom_price_penalty, reward_scaling_quo, reward_lambda}; LobStepConfig.{tick_size, episode_time,
 * n_agent_types, n_messages, agent}; LobStepBuffers.{best_asks, mid_price, agent_f32, perm, message_data, obs, done_all,
 * info_agent_f32, work_redo_count}; LobReplayBuffers.{start, n_msgs, best_out, cancel_u}.  Writes min(n, 25) values,
 * returns 25. */
#define LOB_ABI_N_OFFSETS 25
int32_t lob_abi_offsets(int64_t* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* LOBSTEP_H_ */
